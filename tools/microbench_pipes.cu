// microbench_pipes.cu -- do integer min/max (VIMNMX / VIMNMX3, the DPX family) and fp32 min/max (FMNMX / FMNMX3) share
// an execution pipe on B200?  If they did not, the E reduction of the MSV kernel could move to the integer side.
// Each probe runs 148 CTAs x 1024 threads and reports warp-instructions per clock per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int ITERS = 2048;
constexpr int CH = 16;

template <int KIND> __global__ void __launch_bounds__(1024, 1) probe(int* out, long long* cycles, int seed) {
    float f[CH], g[CH];
    int a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        f[i] = seed + threadIdx.x + i;
        g[i] = seed * 0.5f + i;
        a[i] = seed + threadIdx.x * 3 + i;
        b[i] = seed * 7 + i;
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if constexpr (KIND == 0) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
            if constexpr (KIND == 1) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if constexpr (KIND == 2) a[i] = __vimax3_s32(a[i], b[i], it);
            if constexpr (KIND == 3) { // interleaved fp32 max + int max3: additive if the pipes are distinct
                asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
                a[i] = __vimax3_s32(a[i], b[i], it);
            }
            if constexpr (KIND == 4) { // fp32 max + fp32 add (known distinct pipes) as the reference for "additive"
                asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(f[(i + 1) % CH]));
            }
            if constexpr (KIND == 5) a[i] = __viaddmax_s32(a[i], b[i], it);
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i] + b[i] + static_cast<int>(f[i] + g[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND> double run(int sms, int* out, long long* d_cycles, double instr_per_iter) {
    for (int r = 0; r < 2; ++r) {
        probe<KIND><<<sms, 1024>>>(out, d_cycles, 1);
        if (cudaDeviceSynchronize() != cudaSuccess) exit(1);
    }
    std::vector<long long> cyc(sms);
    cudaMemcpy(cyc.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    long long worst = 0;
    for (auto c : cyc) worst = c > worst ? c : worst;
    return instr_per_iter * ITERS * 32.0 / static_cast<double>(worst); // warp-instructions per clock per SM
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    int* out;
    long long* d_cycles;
    cudaMalloc(&out, sizeof(int) * sms * 1024);
    cudaMalloc(&d_cycles, sizeof(long long) * sms);
    printf("{\"unit\": \"warp-instructions per clock per SM (1024 threads/SM)\"");
    printf(", \"fmnmx\": %.2f", run<0>(sms, out, d_cycles, CH));
    printf(", \"imnmx_s32\": %.2f", run<1>(sms, out, d_cycles, CH));
    printf(", \"vimnmx3_s32\": %.2f", run<2>(sms, out, d_cycles, CH));
    printf(", \"fmnmx_plus_vimnmx3_total\": %.2f", run<3>(sms, out, d_cycles, 2 * CH));
    printf(", \"fmnmx_plus_fadd_total\": %.2f", run<4>(sms, out, d_cycles, 2 * CH));
    printf(", \"viaddmnmx_s32\": %.2f", run<5>(sms, out, d_cycles, CH));
    printf("}\n");
    return 0;
}
