// Microbenchmark: cost of LDS.128 when the 32 lanes of a warp read only a few DISTINCT 16-byte addresses
// (lanes of one cell read the same candidate; a warp of 32 consecutive particles covers ~4 cells).
// Question: is a k-address LDS.128 one shared-memory wavefront (broadcast) or four (quarter-warp phases)?
// Reports cycles per warp-level LDS.128 on one SM with 16 resident warps (4 per scheduler).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const int *__restrict__ lane_index, float *out, int iters, long long *cycles){
    __shared__ float4 s[2048];
    for(int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = lane_index[lane];
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for(int it = 0; it < iters; it++){
#pragma unroll
        for(int u = 0; u < 16; u++){
            float4 v = s[(idx + u * 8) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx = (idx + 1) & 2047;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main(){
    float *o; cudaMalloc(&o, 512 * 4); int *li; cudaMalloc(&li, 32 * 4); long long *cy; cudaMalloc(&cy, 8);
    const char *names[] = {"all lanes one address", "4 groups of 8 (quarter-warp aligned), +1 apart", "3 unaligned groups (11/11/10), +3 apart",
        "5 groups (7/6/7/6/6), +11 apart", "32 distinct consecutive (conflict-free)", "2 groups 128 B apart (same banks)", "4 groups of 8, +8 apart (same banks)",
        "8 groups of 4, +5 apart"};
    for(int pat = 0; pat < 8; pat++){
        int h[32];
        for(int l = 0; l < 32; l++){
            switch(pat){
            case 0: h[l] = 0; break;
            case 1: h[l] = l / 8; break;
            case 2: h[l] = (l < 11 ? 0 : l < 22 ? 1 : 2) * 3; break;
            case 3: h[l] = (l < 7 ? 0 : l < 13 ? 1 : l < 20 ? 2 : l < 26 ? 3 : 4) * 11; break;
            case 4: h[l] = l; break;
            case 5: h[l] = (l / 16) * 8; break;
            case 6: h[l] = (l / 8) * 8; break;
            case 7: h[l] = (l / 4) * 5; break;
            }
        }
        cudaMemcpy(li, h, sizeof(h), cudaMemcpyHostToDevice);
        int iters = 4000; long long c = 0;
        for(int rep = 0; rep < 2; rep++){ k<<<1, 512>>>(li, o, iters, cy); cudaDeviceSynchronize(); }
        cudaMemcpy(&c, cy, 8, cudaMemcpyDeviceToHost);
        // 16 warps x iters x 16 loads through one SM's shared-memory pipe
        printf("%-52s %.2f cycles per warp-level LDS.128 (SM-wide)\n", names[pat], (double)c / (16.0 * iters * 16.0));
    }
    return 0;
}
