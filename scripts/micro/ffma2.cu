// Microbenchmark: FP32 FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template<int MODE> __global__ void k(float *out, int iters){
    float a[8], b = 1.0001f, c = 0.5f;
    unsigned long long p[8], pb, pc;
    for(int i = 0; i < 8; i++){ a[i] = threadIdx.x + i; float2 t = make_float2(a[i], a[i] + 1); p[i] = *reinterpret_cast<unsigned long long *>(&t); }
    { float2 t = make_float2(b, b); pb = *reinterpret_cast<unsigned long long *>(&t); t = make_float2(c, c); pc = *reinterpret_cast<unsigned long long *>(&t); }
    for(int it = 0; it < iters; it++){
#pragma unroll
        for(int i = 0; i < 8; i++){
            if(MODE == 0) a[i] = fmaf(a[i], b, c);
            else asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(pb), "l"(pc));
        }
    }
    float s = 0; for(int i = 0; i < 8; i++){ s += a[i]; float2 t = *reinterpret_cast<float2 *>(&p[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main(){
    float *o; cudaMalloc(&o, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for(int mode = 0; mode < 2; mode++){
        for(int rep = 0; rep < 2; rep++){
            cudaEventRecord(e0);
            if(mode == 0) k<0><<<148 * 8, 256>>>(o, iters); else k<1><<<148 * 8, 256>>>(o, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double inst = 148.0 * 8 * 256 * 8.0 * iters;  // thread-level instructions
            printf("%s: %.3f ms, %.2f T thread-instr/s, %.2f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, inst / ms / 1e9, inst * (mode ? 4 : 2) / ms / 1e9);
        }
    }
    return 0;
}
