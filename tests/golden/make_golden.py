"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/bbref, built from /root/reference by oracle/build_ref.sh) in this container.
The reference cannot travel to the GPU box, the vectors can.  Re-run: python tests/golden/make_golden.py
"""
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
import scenes  # noqa: E402

I = O.mat_str(np.eye(4))


def T(x, y, z):
    return O.mat_str(O.translate(x, y, z))


def load_all(wd, prefix, names):
    return {n: np.load(os.path.join(wd, prefix + n + ".npy")) for n in names}


TRACE = ["rebuild_flag", "cell_count", "cell_order", "nbr_count", "nbr_ids", "density", "eos_pressure", "force_np", "delta", "pos_pred",
         "density_pred", "pressure", "force_p", "max_density_error", "pos_out", "vel_out", "force_out", "rebuild_flag_out"]


def probe_trace():
    """SURVEY D.2 probe scene, particles emitted by the reference's own emitter (glibc rand, seed 1):
    state after 30 sub-steps, its chains, and one traced sub-step from there."""
    job = ["threads 4", "spacing 0.02", "scale 1.8", f"collider box {I} 0.6 0.6 0.6 1 0", "domain_from_collider 0",
           f"emit_box {T(0.1, -0.1, 0.1)} 0.2 0.3 0.2 0 -1 0 0.001 1", "setup", "delta 7e-4",
           "dump {wd}/s0_", "dump_grid {wd}/s0_", "step 7e-4 30", "dump {wd}/s30_", "dump_grid {wd}/s30_",
           "trace 7e-4 {wd}/t_", "step 7e-4 69", "dump {wd}/s100_"]
    out, wd = O.run_ref(job)
    m = re.search(r"mass=(\S+) h=(\S+)", out)
    g = re.search(r"grid min=\((\S+) (\S+) (\S+)\) len=\((\S+) (\S+) (\S+)\)", out)
    c = re.search(r"cells=(\d+) \((\d+) x (\d+) x (\d+)\)", out)
    d = re.search(r"delta\(\S+\)=(\S+)", out)
    data = dict(mass=float(m.group(1)), h=float(m.group(2)), grid_min=[float(g.group(k)) for k in (1, 2, 3)],
                grid_len=[float(g.group(k)) for k in (4, 5, 6)], grid_n=[int(c.group(k)) for k in (2, 3, 4)],
                delta_7e4=float(d.group(1)))
    for pre, names in (("s0_", ["pos", "vel", "cell_count", "cell_order"]),
                       ("s30_", ["pos", "vel", "force", "density", "cell_count", "cell_order", "nbr_count"]),
                       ("s100_", ["pos", "vel"])):
        for k, v in load_all(wd, pre, names).items():
            data[pre + k] = v
    for k, v in load_all(wd, "t_", TRACE).items():
        data["t_" + k] = v
    np.savez_compressed(os.path.join(HERE, "probe_trace.npz"), **data)
    print("probe_trace", len(data["s0_pos"]), "particles")


def collider_vectors():
    """ColliderSet3::ResolveCollision on random (position, velocity) pairs for each collider family."""
    rng = np.random.default_rng(11)
    n = 4000
    pos = rng.uniform(-0.36, 0.36, size=(n, 3))
    vel = rng.normal(0, 2.0, size=(n, 3))
    torus = scenes.sdf_torus((0.05, -0.1, 0.0), 0.12, 0.04)
    sets = {
        "container_box": [f"collider box {I} 0.6 0.6 0.6 1 0"],
        "container_sphere": [f"collider sphere {I} 0.3 1 0.3"],
        "box_and_sphere": [f"collider box {I} 0.7 0.7 0.7 1 0", f"collider box {T(0.1, -0.2, 0.0)} 0.2 0.1 0.3 0 0.25",
                           f"collider sphere {T(-0.1, 0.1, 0.1)} 0.09 0 0.5"],
        "sdf_torus": [f"collider box {I} 0.7 0.7 0.7 1 0", "SDF"],
    }
    data = dict(pos=pos, vel=vel)
    for name, lines in sets.items():
        wd = tempfile.mkdtemp(prefix="bbref_")
        O.write_particles(os.path.join(wd, "q.bin"), pos, vel)
        job = ["threads 1", "spacing 0.02", "scale 1.8"]
        for l in lines:
            if l == "SDF":
                bmin, bmax = (-0.15, -0.16, -0.2), (0.25, -0.04, 0.2)
                import bubbles_b200 as bb
                nodes, dx, origin = bb.sdf_grid_layout(bmin, bmax, 0.01, 0.1)
                ix, iy, iz = np.meshgrid(np.arange(nodes[0]), np.arange(nodes[1]), np.arange(nodes[2]), indexing="ij")
                pts = np.stack([origin[0] + dx * ix, origin[1] + dx * iy, origin[2] + dx * iz], axis=-1)
                field = np.ascontiguousarray(torus(pts.reshape(-1, 3)).reshape(nodes).transpose(2, 1, 0))
                field.tofile(os.path.join(wd, "sdf.bin"))
                job.append(f"collider sdf {bmin[0]} {bmin[1]} {bmin[2]} {bmax[0]} {bmax[1]} {bmax[2]} 0.01 0.1 0.1 {wd}/sdf.bin")
            else:
                job.append(l)
        job += ["domain -0.4 -0.4 -0.4 0.4 0.4 0.4", f"particles {wd}/q.bin", "setup"]
        for tag, rad, rest in (("a", 0.02, 0.0), ("b", 0.02, 0.6)):
            job.append(f"collide {wd}/q.bin {rad} {rest} {wd}/{name}_{tag}_")
        O.run_ref(job, wd)
        for tag in ("a", "b"):
            for k in ("pos", "vel", "hit"):
                data[f"{name}_{tag}_{k}"] = np.load(os.path.join(wd, f"{name}_{tag}_{k}.npy"))
        print(name, "hits", int(data[f"{name}_a_hit"].sum()), int(data[f"{name}_b_hit"].sum()))
    # a moving, spinning sphere obstacle (Shape::SetVelocities -> VelocityAt in the response) and a collider switched off
    wd = tempfile.mkdtemp(prefix="bbref_")
    O.write_particles(os.path.join(wd, "q.bin"), pos, vel)
    job = ["threads 1", "spacing 0.02", "scale 1.8", f"collider box {I} 0.7 0.7 0.7 1 0",
           f"collider sphere {T(-0.1, 0.1, 0.1)} 0.12 0 0.5", f"collider box {T(0.1, -0.2, 0.0)} 0.2 0.1 0.3 0 0.25",
           "domain -0.4 -0.4 -0.4 0.4 0.4 0.4", f"particles {wd}/q.bin", "setup",
           "collider_velocity 1 0.5 -0.25 1.5 0 3 -2", f"collide {wd}/q.bin 0.02 0.6 {wd}/mv_",
           "collider_active 2 0", f"collide {wd}/q.bin 0.02 0.6 {wd}/off_"]
    O.run_ref(job, wd)
    for tag in ("mv", "off"):
        for k in ("pos", "vel", "hit"):
            data[f"moving_{tag}_{k}"] = np.load(os.path.join(wd, f"{tag}_{k}.npy"))
    print("moving sphere hits", int(data["moving_mv_hit"].sum()), "with the box off", int(data["moving_off_hit"].sum()))
    np.savez_compressed(os.path.join(HERE, "collider_vectors.npz"), **data)


def obstacle_run():
    """Block dropped on a sphere + box obstacle, 240 sub-steps, fixed dt and then two CFL frames (Advance)."""
    job = ["threads 4", "spacing 0.02", "scale 1.8", f"collider box {I} 0.6 0.6 0.6 1 0",
           f"collider sphere {T(0.1, -0.27, 0.1)} 0.08 0 0.2", f"collider box {T(-0.15, -0.25, -0.1)} 0.1 0.1 0.1 0 0.1",
           "domain_from_collider 0", f"emit_box {T(0.05, -0.1, 0.05)} 0.24 0.3 0.24 0 -2 0 0.001 7", "setup",
           "dump {wd}/s0_", "step 7e-4 240", "dump {wd}/s240_", "dump_grid {wd}/s240_",
           "advance 0.004166666666666667", "advance 0.004166666666666667", "dump {wd}/adv_", "trace 7e-4 {wd}/t_"]
    out, wd = O.run_ref(job)
    data = {}
    for k, v in load_all(wd, "t_", ["density", "eos_pressure", "nbr_count", "pos_out"]).items():  # a compressed state: Tait EOS > 0
        data["t_" + k] = v
    for pre, names in (("s0_", ["pos", "vel"]), ("s240_", ["pos", "vel", "density", "cell_count", "cell_order", "nbr_count"]),
                       ("adv_", ["pos", "vel"])):
        for k, v in load_all(wd, pre, names).items():
            data[pre + k] = v
    np.savez_compressed(os.path.join(HERE, "obstacle_run.npz"), **data)
    print("obstacle_run", len(data["s0_pos"]), "particles")


def append_block():
    """The particles appended in append_run: a small jitter-free BCC block above the probe scene's fluid, moving down."""
    pts = O.bcc_points((-0.2, 0.12, -0.2), (-0.08, 0.2, -0.06), 0.02)
    return pts, np.tile(np.array([0.5, -2.0, 0.25]), (len(pts), 1))


def append_run():
    """Continuous emission (ContinuousParticleSetBuilder3::AddParticle + Commit, src/core/grid.h:1409-1441): probe scene,
    20 sub-steps, append a block, chains right after the append, one traced sub-step, 30 more sub-steps, a second
    append (the same block again, shifted), 20 more sub-steps."""
    sc = scenes.probe_scene()
    wd = tempfile.mkdtemp(prefix="bbref_")
    O.write_particles(os.path.join(wd, "p.bin"), sc["pos"], sc["vel"])
    add_pos, add_vel = append_block()
    O.write_particles(os.path.join(wd, "a1.bin"), add_pos, add_vel)
    O.write_particles(os.path.join(wd, "a2.bin"), add_pos + np.array([0.3, 0.0, 0.2]), add_vel)
    job = ["threads 4", "spacing 0.02", "scale 1.8", f"collider box {I} 0.6 0.6 0.6 1 0", "domain_from_collider 0",
           "continuous 6000", f"particles {wd}/p.bin", "setup", "step 7e-4 20", "dump {wd}/s20_", "dump_grid {wd}/s20_",
           f"append {wd}/a1.bin", "dump_grid {wd}/a1_", "trace 7e-4 {wd}/t_", "step 7e-4 30",
           f"append {wd}/a2.bin", "dump_grid {wd}/a2_", "step 7e-4 20", "dump {wd}/end_", "dump_grid {wd}/end_"]
    out, _ = O.run_ref(job, wd)
    data = dict(p_pos=sc["pos"].astype(np.float64), p_vel=sc["vel"].astype(np.float64), add_pos=add_pos, add_vel=add_vel)
    for pre, names in (("s20_", ["pos", "vel", "cell_count", "cell_order"]), ("a1_", ["cell_count", "cell_order"]),
                       ("a2_", ["cell_count", "cell_order"]), ("end_", ["pos", "vel", "density", "cell_count", "cell_order"])):
        for k, v in load_all(wd, pre, names).items():
            data[pre + k] = v
    for k, v in load_all(wd, "t_", TRACE).items():
        data["t_" + k] = v
    np.savez_compressed(os.path.join(HERE, "append_run.npz"), **data)
    print("append_run", len(sc["pos"]), "+", len(add_pos), "+", len(add_pos), "particles")


def emit_run():
    """Continuous re-emission (ContinuousParticleSetBuilder3::MapGrid at setup + MapGridEmit between steps,
    src/core/grid.h:1288-1407): the probe block falls out of its initial cells, the mapped cells re-emit twice."""
    sc = scenes.probe_scene()
    wd = tempfile.mkdtemp(prefix="bbref_")
    O.write_particles(os.path.join(wd, "p.bin"), sc["pos"], sc["vel"])
    job = ["threads 4", "spacing 0.02", "scale 1.8", f"collider box {I} 0.6 0.6 0.6 1 0", "domain_from_collider 0",
           "continuous 12000", f"particles {wd}/p.bin", "setup", "dump_grid {wd}/s0_", "step 7e-4 40", "dump {wd}/s40_", "dump_grid {wd}/s40_",
           "map_emit 0 -1 0 0.02", "dump {wd}/e1_", "dump_grid {wd}/e1_", "step 7e-4 30", "map_emit 0 -1 0 0.02",
           "dump {wd}/e2_", "dump_grid {wd}/e2_", "step 7e-4 10", "dump {wd}/end_", "dump_grid {wd}/end_"]
    out, _ = O.run_ref(job, wd)
    added = [int(x) for x in re.findall(r"map_emit added=(\d+)", out)]
    data = dict(p_pos=sc["pos"].astype(np.float64), p_vel=sc["vel"].astype(np.float64), added=np.array(added))
    for pre, names in (("s0_", ["cell_count", "cell_order"]), ("s40_", ["pos", "vel", "cell_count", "cell_order"]),
                       ("e1_", ["pos", "vel", "cell_count", "cell_order"]), ("e2_", ["pos", "vel", "cell_count", "cell_order"]),
                       ("end_", ["pos", "vel", "cell_order"])):
        for k, v in load_all(wd, pre, names).items():
            data[pre + k] = v
    np.savez_compressed(os.path.join(HERE, "emit_run.npz"), **data)
    print("emit_run", len(sc["pos"]), "particles, added", added)


def pseudo_run():
    """Pseudo-viscosity smoothing switched on (coefficient 200: 200 * 7e-4 > 0.1; ComputePseudoViscosity*KernelFor,
    src/equations/sph_equations3.cpp:341-382, 469-483): probe scene, 25 sub-steps."""
    sc = scenes.probe_scene()
    wd = tempfile.mkdtemp(prefix="bbref_")
    O.write_particles(os.path.join(wd, "p.bin"), sc["pos"], sc["vel"])
    job = ["threads 4", "spacing 0.02", "scale 1.8", f"collider box {I} 0.6 0.6 0.6 1 0", "domain_from_collider 0",
           f"particles {wd}/p.bin", "setup", "pseudo 200", "step 7e-4 1", "dump {wd}/s1_", "step 7e-4 24", "dump {wd}/s25_"]
    O.run_ref(job, wd)
    data = dict(p_pos=sc["pos"].astype(np.float64), p_vel=sc["vel"].astype(np.float64))
    for pre in ("s1_", "s25_"):
        for k, v in load_all(wd, pre, ["pos", "vel", "density"]).items():
            data[pre + k] = v
    np.savez_compressed(os.path.join(HERE, "pseudo_run.npz"), **data)
    print("pseudo_run", len(sc["pos"]), "particles")


def sdf_run():
    """Block dropped on a baked-SDF torus (MakeSDFShape-style vertex grid, Shape::ClosestPointBySDF in the response),
    150 sub-steps: the SDF collider inside a whole trajectory."""
    import bubbles_b200 as bb
    torus = scenes.sdf_torus((0.05, -0.1, 0.0), 0.12, 0.04)
    bmin, bmax = (-0.15, -0.16, -0.2), (0.25, -0.04, 0.2)
    nodes, dx, origin = bb.sdf_grid_layout(bmin, bmax, 0.01, 0.1)
    ix, iy, iz = np.meshgrid(np.arange(nodes[0]), np.arange(nodes[1]), np.arange(nodes[2]), indexing="ij")
    pts = np.stack([origin[0] + dx * ix, origin[1] + dx * iy, origin[2] + dx * iz], axis=-1)
    field = np.ascontiguousarray(torus(pts.reshape(-1, 3)).reshape(nodes).transpose(2, 1, 0))
    wd = tempfile.mkdtemp(prefix="bbref_")
    field.tofile(os.path.join(wd, "sdf.bin"))
    job = ["threads 4", "spacing 0.02", "scale 1.8", f"collider box {I} 0.7 0.7 0.7 1 0",
           f"collider sdf {bmin[0]} {bmin[1]} {bmin[2]} {bmax[0]} {bmax[1]} {bmax[2]} 0.01 0.1 0.1 {wd}/sdf.bin",
           "domain_from_collider 0", f"emit_box {T(0.05, 0.12, 0.0)} 0.2 0.2 0.2 0 -2.5 0 0.001 3", "setup",
           "dump {wd}/s0_", "step 7e-4 150", "dump {wd}/s150_", "dump_grid {wd}/s150_"]
    O.run_ref(job, wd)
    data = {}
    for pre, names in (("s0_", ["pos", "vel"]), ("s150_", ["pos", "vel", "density", "cell_count", "cell_order"])):
        for k, v in load_all(wd, pre, names).items():
            data[pre + k] = v
    np.savez_compressed(os.path.join(HERE, "sdf_run.npz"), **data)
    print("sdf_run", len(data["s0_pos"]), "particles; moved by the torus:",
          int((np.abs(data["s150_vel"][:, 0]) > 1e-3).sum()))


def sph_run():
    """SphSolver3 (BASELINE configs[0]'s solver, src/solvers/sph_solver3.cpp:48-67) on ONE thread -- the only deterministic
    mode of its CPU path: ComputeAllForcesFor integrates every particle inside its own force evaluation
    (src/equations/sph_equations3.cpp:184-278), so with one thread particle i sees the moved neighbours j < i
    (Gauss-Seidel).  Probe scene with a sphere obstacle on the floor, fixed dt = 0.4 h / c_s = 1.44e-4: state after 1,
    20 and 120 sub-steps (the block reaches the floor and the sphere), chains and lists of sub-step 1."""
    job = ["solver sph", "threads 1", "spacing 0.02", "scale 1.8", f"collider box {I} 0.6 0.6 0.6 1 0",
           f"collider sphere {T(0.1, -0.27, 0.1)} 0.08 0 0.2", "domain_from_collider 0",
           f"emit_box {T(0.1, -0.02, 0.1)} 0.2 0.3 0.2 0 -6 0 0.001 1", "setup", "dump {wd}/s0_",
           "step 1.44e-4 1", "dump {wd}/s1_", "dump_grid {wd}/s1_", "step 1.44e-4 19", "dump {wd}/s20_",
           "step 1.44e-4 100", "dump {wd}/s120_", "dump_grid {wd}/s120_"]
    out, wd = O.run_ref(job)
    data = {}
    for pre, names in (("s0_", ["pos", "vel"]), ("s1_", ["pos", "vel", "force", "density", "pressure", "cell_count", "cell_order", "nbr_count"]),
                       ("s20_", ["pos", "vel", "force", "density", "pressure"]),
                       ("s120_", ["pos", "vel", "force", "density", "pressure", "cell_count", "cell_order", "nbr_count"])):
        for k, v in load_all(wd, pre, names).items():
            data[pre + k] = v
    np.savez_compressed(os.path.join(HERE, "sph_run.npz"), **data)
    print("sph_run", len(data["s0_pos"]), "particles")


def mesh_collider():
    """Triangle-mesh collider (MakeMesh + the SDF grid GenerateShapeSDF builds for it: BVH closest distance, sign by ray
    parity -- src/shapes/bvh.cpp:51-56, 500-557, src/core/shape.cpp:358-377, 411-431, 479-511), on a procedural closed torus
    mesh: the grid the reference generates, Shape::ClosestDistance and the collider-set response at 4 000 random points, and
    a 150-sub-step run of the probe block dropped on it."""
    P, Tr = scenes.torus_mesh((0.1, -0.26, 0.1), 0.09, 0.03)
    wd = tempfile.mkdtemp(prefix="bbref_")
    with open(os.path.join(wd, "m.bin"), "wb") as f:
        f.write(np.int64(len(P)).tobytes()); f.write(np.int64(len(Tr)).tobytes()); f.write(P.tobytes()); f.write(Tr.tobytes())
    rng = np.random.default_rng(11)
    q = rng.uniform([-0.05, -0.31, -0.05], [0.25, -0.2, 0.25], size=(4000, 3))
    v = rng.normal(0, 1.5, size=(4000, 3))
    O.write_particles(os.path.join(wd, "q.bin"), q, v)
    sc = scenes.probe_scene()
    O.write_particles(os.path.join(wd, "p.bin"), sc["pos"], sc["vel"])
    job = ["threads 4", "spacing 0.02", "scale 1.8", f"collider box {I} 0.6 0.6 0.6 1 0", f"collider mesh {wd}/m.bin 0 0.1 0.01 0.1",
           "domain_from_collider 0", f"particles {wd}/p.bin", "setup", f"dump_sdf 1 {wd}/sdf_", f"closest_distance 1 {wd}/q.bin {wd}/cd_",
           f"collide {wd}/q.bin 0.02 0 {wd}/a_", f"collide {wd}/q.bin 0.02 0.6 {wd}/b_", "step 7e-4 150", "dump {wd}/s150_", "dump_grid {wd}/s150_"]
    O.run_ref(job, wd)
    data = dict(vertices=P, triangles=Tr, q_pos=q, q_vel=v, p_pos=sc["pos"].astype(np.float64), p_vel=sc["vel"].astype(np.float64))
    for k in ("res", "meta", "field", "bounds"):
        data["sdf_" + k] = np.load(os.path.join(wd, f"sdf_{k}.npy"))
    data["distance"] = np.load(os.path.join(wd, "cd_distance.npy"))
    for tag in ("a", "b"):
        for k in ("pos", "vel", "hit"):
            data[f"{tag}_{k}"] = np.load(os.path.join(wd, f"{tag}_{k}.npy"))
    for k, val in load_all(wd, "s150_", ["pos", "vel", "cell_count", "cell_order"]).items():
        data["s150_" + k] = val
    np.savez_compressed(os.path.join(HERE, "mesh_collider.npz"), **data)
    print("mesh_collider", len(Tr), "triangles, sdf", data["sdf_res"], "hits", int(data["a_hit"].sum()), int(data["b_hit"].sum()),
          "inside nodes", int((data["sdf_field"] < 0).sum()))


def grid_facts():
    """UtilBuildGridForDomain results printed by the reference for several domains / spacings."""
    rows = []
    for (lo, hi, s, k) in [((-0.3, -0.3, -0.3), (0.3, 0.3, 0.3), 0.02, 1.8), ((-1.625, -1.5, -1.625), (1.625, 1.5, 1.625), 0.02, 1.8),
                           ((0, 0, 0), (1, 2, 3), 0.05, 2.0), ((-1, -1, -1), (1, 1, 1), 0.1, 2.0), ((-0.7, 0.1, 2.0), (0.9, 1.3, 2.55), 0.013, 1.7)]:
        wd = tempfile.mkdtemp(prefix="bbref_")
        O.write_particles(os.path.join(wd, "q.bin"), np.array([[0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])]]), np.zeros((1, 3)))
        job = ["threads 1", f"spacing {s}", f"scale {k}", f"collider box {I} 10 10 10 1 0",
               f"domain {lo[0]} {lo[1]} {lo[2]} {hi[0]} {hi[1]} {hi[2]}", f"particles {wd}/q.bin", "setup"]
        out, _ = O.run_ref(job, wd)
        g = re.search(r"grid min=\((\S+) (\S+) (\S+)\) len=\((\S+) (\S+) (\S+)\)", out)
        c = re.search(r"cells=(\d+) \((\d+) x (\d+) x (\d+)\)", out)
        m = re.search(r"mass=(\S+) h=(\S+)", out)
        rows.append([*lo, *hi, s, k, *[float(g.group(i)) for i in range(1, 7)], *[int(c.group(i)) for i in (2, 3, 4)], float(m.group(1))])
    np.save(os.path.join(HERE, "grid_facts.npy"), np.array(rows))
    print("grid_facts", len(rows))


TSEQ_JOB = [  # keyframes K = Translate(t) * Rotate(angle [deg], axis) * Scale(s): t angle axis s | t angle axis s | s0 s1
    "tseq_add 0 0 0 0 0 1 0 1   0.3 0.1 -0.2 0 0 1 0 1   0 1",                      # pure translation
    "tseq_add 0.3 0.1 -0.2 0 0 1 0 1   0.3 0.1 -0.2 170 1 1 0 1   1 2",             # rotation up to 170 deg about (1, 1, 0)
    "tseq_add 0.3 0.1 -0.2 170 1 1 0 1   0.5 0.4 0.0 350 1 1 0 1.5   2 3.5",        # through 180 deg (trace <= 0, quaternion flip) while scaling
    "tseq_add 0.5 0.4 0.0 350 1 1 0 1.5   -0.2 0.0 0.1 20 0.3 -1 0.5 0.75   3.5 4",  # new axis, shrink
    "tseq_restore 4 5",
    "tseq_eval -0.25 0.0625 90 {wd}/t_",
    "qseq_add 0 0 0 1 0", "qseq_add 120 0 1 1 1", "qseq_add 300 1 0 0 2", "qseq_add 181 0.2 0.3 -1 2.75",
    "qseq_eval -0.5 0.125 32 {wd}/q_",
]


def transform_sequence():
    """TransformSequence / QuaternionSequence of the unmodified reference (src/core/transform_sequence.cpp) on a keyframe
    script: interpolated matrices, their inverses, and the linear / angular velocities it reports call by call."""
    wd = tempfile.mkdtemp(prefix="bbref_")
    O.run_ref([l.format(wd=wd) for l in TSEQ_JOB], wd)
    data = {}
    for pre in ("t_", "q_"):
        data.update({pre + k: v for k, v in load_all(wd, pre, ["m", "minv", "linear", "angular"]).items()})
    data["job"] = np.array(TSEQ_JOB)
    np.savez_compressed(os.path.join(HERE, "transform_sequence.npz"), **data)
    print("transform_sequence", data["t_m"].shape, data["q_m"].shape)


if __name__ == "__main__":
    assert O.ref_available(), "run oracle/build_ref.sh first"
    if "--only-transform-sequence" in sys.argv:
        transform_sequence()
        sys.exit(0)
    transform_sequence()
    probe_trace()
    sph_run()
    mesh_collider()
    collider_vectors()
    obstacle_run()
    append_run()
    emit_run()
    pseudo_run()
    sdf_run()
    grid_facts()
