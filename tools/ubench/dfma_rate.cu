// Micro-benchmark: issue rates of the instructions the extended-precision gather leans on (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/dfma_rate tools/ubench/dfma_rate.cu
// Prints results per SM per clock for DFMA, FFMA, F2F.F64.F32, IMAD.MOV-free integer ops (LOP3 / SHL) and a DFMA + 1 ALU mix.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3;
    const float fa = (float)a, fb = (float)b;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // 8 independent DFMA chains
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        } else if (MODE == 1) {  // FFMA
            f0 = fmaf(f0, fa, fb); f1 = fmaf(f1, fa, fb); f2 = fmaf(f2, fa, fb); f3 = fmaf(f3, fa, fb);
            f4 = fmaf(f4, fa, fb); f5 = fmaf(f5, fa, fb); f6 = fmaf(f6, fa, fb); f7 = fmaf(f7, fa, fb);
        } else if (MODE == 2) {  // F2F.F64.F32 + DADD to keep it alive
            x0 += (double)f0; x1 += (double)f1; x2 += (double)f2; x3 += (double)f3;
            f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f;
        } else if (MODE == 3) {  // DFMA with an operand assembled from two 32-bit registers (1 ALU op each)
            x0 = fma(__hiloint2double(__double2hiint(x0), u0 << 16), a, x0);
            x1 = fma(__hiloint2double(__double2hiint(x1), u1 & 0xffff0000u), a, x1);
            x2 = fma(__hiloint2double(__double2hiint(x2), u2 << 16), a, x2);
            x3 = fma(__hiloint2double(__double2hiint(x3), u3 & 0xffff0000u), a, x3);
            u0 += 3; u1 += 5; u2 += 7; u3 += 11;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + u0 + u1 + u2 + u3;
}

template <int MODE>
void run(const char *name, int per_iter) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    k<MODE><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ops = (double)blocks * threads * iters * per_iter;
    printf("%-28s %8.3f ms  %.3e ops/s  %.1f per SM per clk (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3),
           ops / (ms * 1e-3) / p.multiProcessorCount / (khz * 1e3), khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("DFMA", 8);
    run<1>("FFMA", 8);
    run<2>("F2F.F64.F32 (+DADD,FADD)", 4);
    run<3>("DFMA + pair assembly", 4);
    return 0;
}
