// fp64_peak.cu -- DFMA micro-benchmark: the FP64-pipe roof this pool's B200s actually reach
// (MEASURED_PEAKS.json has no FP64 entry).  Prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/fp64_peak tools/fp64_peak.cu
// (-fmad=false keeps the DFMA/DMUL/DADD mix of dmix as written; the fma() calls stay DFMAs)
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) dfma(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x*1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

// the same with THREE register operands per DFMA (what a real stencil kernel issues: 6 register
// reads per instruction instead of 2), and a DFMA/DMUL/DADD mix in the proportions of the sweep
// kernels (5 : 4 : 2)
template <int ILP>
__global__ void __launch_bounds__(256) dfma3(double *out, int iters, double a, double b) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x*1e-3 + i; y[i] = a + i*1e-9; z[i] = b + i*1e-12; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], y[(i + 1) % ILP], z[(i + 3) % ILP]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + y[i] + z[i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) dmix(double *out, int iters, double a, double b) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x*1e-3 + i; y[i] = a + i*1e-9; z[i] = b + i*1e-12; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {        // 11 instructions: 5 DFMA, 4 DMUL, 2 DADD
                x[i] = fma(x[i], y[(i + 1) % ILP], z[(i + 3) % ILP]);
                x[i] = x[i]*y[i];
                x[i] = fma(x[i], z[(i + 1) % ILP], y[(i + 2) % ILP]);
                x[i] = x[i] + z[i];
                x[i] = x[i]*y[(i + 5) % ILP];
                x[i] = fma(y[i], z[(i + 2) % ILP], x[i]);
                x[i] = x[i]*z[(i + 5) % ILP];
                x[i] = fma(x[i], y[(i + 3) % ILP], z[(i + 6) % ILP]);
                x[i] = x[i] + y[(i + 6) % ILP];
                x[i] = x[i]*z[(i + 7) % ILP];
                x[i] = fma(x[i], y[(i + 7) % ILP], z[(i + 4) % ILP]);
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + y[i] + z[i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

__global__ void ddiv(double *out, int iters, double a) {
    double x = 1.0 + threadIdx.x*1e-3, y = 2.0 + threadIdx.x*1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) { x = a/x; y = a/y; }
    }
    out[blockIdx.x*blockDim.x + threadIdx.x] = x + y;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount*8, threads = 256;
    double *out; cudaMalloc(&out, sizeof(double)*blocks*threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int ILP = 8, iters = 4096;
    double best = 0, sustained = 0;
    dfma<ILP><<<blocks, threads>>>(out, 64, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    for (int rep = 0; rep < 10; rep++) {
        cudaEventRecord(e0);
        dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0*16*ILP*(double)iters*blocks*threads/(ms*1e-3)/1e12;
        if (tf > best) best = tf;
    }
    // sustained: back to back for ~3 s
    cudaEventRecord(e0);
    int n = 0; float total = 0;
    while (total < 3000.f) {
        for (int r = 0; r < 20; r++) dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        n += 20;
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&total, e0, e1);
    }
    sustained = 2.0*16*ILP*(double)iters*blocks*threads*n/(total*1e-3)/1e12;
    // three register operands per DFMA, and the DFMA/DMUL/DADD mix: FP64 instructions per second
    double g3 = 0, gmix = 0;
    dfma3<ILP><<<blocks, threads>>>(out, 64, 1.0000001, 1e-9);
    dmix<ILP><<<blocks, threads>>>(out, 64, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    for (int rep = 0; rep < 5; rep++) {
        float ms;
        cudaEventRecord(e0);
        dfma3<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double t = 16.0*ILP*(double)iters*blocks*threads/(ms*1e-3)/1e12;
        if (t > g3) g3 = t;
        cudaEventRecord(e0);
        dmix<ILP><<<blocks, threads>>>(out, iters/4, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        t = 4.0*11*ILP*(double)(iters/4)*blocks*threads/(ms*1e-3)/1e12;
        if (t > gmix) gmix = t;
    }
    // division throughput (results/s)
    ddiv<<<blocks, threads>>>(out, 16, 3.0); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    ddiv<<<blocks, threads>>>(out, 1024, 3.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double gdiv = 2.0*16*1024.0*blocks*threads/(ms*1e-3)/1e9;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops_burst\": %.2f, \"fp64_dfma_tflops_sustained\": %.2f, "
           "\"fp64_tinst_per_s_dfma_1reg\": %.2f, \"fp64_tinst_per_s_dfma_3reg\": %.2f, \"fp64_tinst_per_s_mix_5fma_4mul_2add\": %.2f, "
           "\"fp64_div_gops\": %.1f, \"dfma_per_div\": %.2f, \"clock_mhz_max\": %d}\n",
           p.name, p.multiProcessorCount, best, sustained, best/2.0, g3, gmix, gdiv, (sustained*1e3/2.0)/gdiv, p.clockRate/1000);
    return 0;
}
