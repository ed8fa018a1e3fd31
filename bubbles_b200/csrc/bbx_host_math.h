// Host-side scalar setup of the engine (FP64): grid sizing, particle mass, PCISPH delta.
// Reference semantics (Bubbles tree): UtilBuildGridForDomain src/core/util.cpp:269-287, Grid::Build
// src/core/grid.h:626-653, SphParticleSet3::ComputeMass src/core/particle.h:557-582,
// PciSphSolver3::ComputeDeltaDenom/ComputeBeta/ComputeDelta src/solvers/pcisph_solver3.cpp:148-189,
// BccLatticePointGenerator src/generator/bcclattice.cpp:5-77.
#pragma once
#include <cmath>
#include <vector>
#include "../../include/bbx.h"

static const double kBbxPi = 3.14159265358979323846;

struct BbxVec3 { double x, y, z; };

// Body-centred-cubic lattice over [lo, hi]: z layers at spacing/2, odd layers shifted by spacing/2.
template<typename F>
static inline void bbxh_bcc_for_each(const double lo[3], const double hi[3], double spacing, F &&fn){
    const double half = spacing / 2;
    const double ext[3] = {std::fabs(hi[0] - lo[0]), std::fabs(hi[1] - lo[1]), std::fabs(hi[2] - lo[2])};
    bool shifted = false;
    for(int k = 0; k * half <= ext[2]; k++){
        const double off = shifted ? half : 0.0;
        const double z = k * half + lo[2];
        for(int j = 0; j * spacing + off <= ext[1]; j++){
            const double y = j * spacing + off + lo[1];
            for(int i = 0; i * spacing + off <= ext[0]; i++){
                if(!fn(BbxVec3{i * spacing + off + lo[0], y, z})) return;
            }
        }
        shifted = !shifted;
    }
}

static inline double bbxh_w_std(double d, double h){
    const double h2 = h * h, d2 = d * d, of = d2 - h2;
    if(std::fabs(of) < 1e-8 || of > 0) return 0.0;
    const double x = 1.0 - d2 / h2;
    return 315.0 / (64.0 * kBbxPi * (h2 * h)) * x * x * x;
}
static inline double bbxh_dw_spiky(double d, double h){
    const double of = d - h;
    if(std::fabs(of) < 1e-8 || of > 0) return 0.0;
    const double h2 = h * h, x = 1.0 - d / h;
    return -45.0 / (kBbxPi * (h2 * h2)) * x * x;
}

// mass such that the densest lattice site of a BCC block at rest has the target density
static inline double bbxh_compute_mass(double h, double spacing, double rho0){
    std::vector<BbxVec3> pts;
    const double lo[3] = {-1.5 * h, -1.5 * h, -1.5 * h}, hi[3] = {1.5 * h, 1.5 * h, 1.5 * h};
    bbxh_bcc_for_each(lo, hi, spacing, [&](const BbxVec3 &p){ if(pts.size() >= 1024) return false; pts.push_back(p); return true; });
    double best = 0.0;
    for(const BbxVec3 &a : pts){
        double sum = 0.0;
        for(const BbxVec3 &b : pts){
            const double x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
            sum += bbxh_w_std(std::sqrt(x * x + y * y + z * z), h);
        }
        if(best < sum) best = sum;
    }
    return rho0 / best;
}

static inline double bbxh_delta_denom(double h, double spacing){
    const double lo[3] = {-1.5 * h, -1.5 * h, -1.5 * h}, hi[3] = {1.5 * h, 1.5 * h, 1.5 * h};
    const double h2 = h * h;
    double gsum[3] = {0, 0, 0}, g2 = 0.0;
    bbxh_bcc_for_each(lo, hi, spacing, [&](const BbxVec3 &p){
        const double d2 = p.x * p.x + p.y * p.y + p.z * p.z;
        if(d2 < h2){
            const double d = std::sqrt(d2);
            double dir[3] = {0, 0, 0};
            if(d > 0){ const double inv = 1.0 / d; dir[0] = p.x * inv; dir[1] = p.y * inv; dir[2] = p.z * inv; }
            const double m = -bbxh_dw_spiky(d, h);
            const double g[3] = {m * dir[0], m * dir[1], m * dir[2]};
            gsum[0] += g[0]; gsum[1] += g[1]; gsum[2] += g[2];
            g2 += g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
        }
        return true;
    });
    double denom = 0.0;
    denom += -(gsum[0] * gsum[0] + gsum[1] * gsum[1] + gsum[2] * gsum[2]) - g2;
    return denom;
}

static inline double bbxh_delta(double mass_over_rho0_sq, double delta_denom, double dt){
    const double beta = 2.0 * mass_over_rho0_sq * (dt * dt);
    return std::fabs(delta_denom) > 0 ? -1 / (beta * delta_denom) : 0.0;
}

static inline void bbxh_grid_build(const int res[3], const double p0[3], const double p1[3], bbx_grid_desc *g){
    g->total = 1;
    for(int k = 0; k < 3; k++){
        const double hi = p0[k] < p1[k] ? p1[k] : p0[k];
        const double lo = p0[k] < p1[k] ? p0[k] : p1[k];
        const double s = hi - lo;
        const double len = s / (double)res[k];
        g->cell_len[k] = len;
        g->n[k] = (int)std::ceil(s / len);
        g->min[k] = lo;
        g->max[k] = lo + (double)g->n[k] * len;
        g->total *= g->n[k];
    }
}

static inline void bbxh_grid_for_domain(const double dmin[3], const double dmax[3], double spacing, double scale, bbx_grid_desc *g){
    const double length = spacing * scale, hlen = 0.5 * length, inv_len = 1.0f / length;
    int res[3]; double p0[3], p1[3];
    for(int k = 0; k < 3; k++){
        int m = (int)std::ceil(std::fabs(dmax[k] - dmin[k]) * inv_len);
        m = (int)((double)m + (m % 2) * length); // the reference's `mx += (mx % 2) * length` (int += double)
        res[k] = m;
        const double center = (dmin[k] + dmax[k]) * 0.5, half = m * hlen;
        p0[k] = center - half; p1[k] = center + half;
    }
    bbxh_grid_build(res, p0, p1, g);
}
