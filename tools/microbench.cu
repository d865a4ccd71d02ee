// microbench.cu -- measures, on the B200 itself, the per-SM issue rates the MSV kernel's roofline depends on
// (SURVEY.md section 8d asks for them): FADD, FMNMX, FMNMX3, the add+max mix of the recurrence, FADD2 (packed fp32x2),
// conflict-free LDS.128, CREDUX.MAX.F32, SHFL, and TMEM reads (tcgen05.ld) as a second on-chip source of emissions.
// Every kernel runs 148 x CTAS_PER_SM blocks; each block times itself with clock64(); the result is
// lane-operations (or bytes) per clock per SM, independent of the SM clock.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu && build/microbench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CHECK(x)                                                                                                       \
    do {                                                                                                               \
        cudaError_t e = (x);                                                                                           \
        if (e != cudaSuccess) {                                                                                        \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                                                    \
            exit(1);                                                                                                   \
        }                                                                                                              \
    } while (0)

constexpr int ITERS = 2048;
constexpr int CHAINS = 16;

enum Kind { K_FADD, K_FMNMX, K_FMNMX3, K_MIX, K_FADD2, K_MIX_FADD2, K_LDS128, K_MIX_LDS, K_CREDUX, K_SHFL, K_COUNT };
const char* kind_name[] = {"fadd", "fmnmx", "fmnmx3", "mix_fadd_fmnmx_half_fmnmx3", "fadd2", "mix_fadd2_fmnmx_half_fmnmx3",
                           "lds128", "mix_with_lds128", "credux_max_f32", "shfl_up"};
// lane-operations counted per inner iteration per thread (for *_lds: bytes read)
template <int KIND> __global__ void __launch_bounds__(1024, 1) probe(float* out, long long* cycles, float seed) {
    extern __shared__ __align__(16) float smem[];
    float a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = seed + threadIdx.x + i;
        b[i] = seed * 0.5f + i;
    }
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) smem[i] = seed + i;
    __syncthreads();
    float e = seed;
    const float4* lane4 = reinterpret_cast<const float4*>(smem) + (threadIdx.x & 31);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if constexpr (KIND == K_FADD) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if constexpr (KIND == K_FMNMX) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if constexpr (KIND == K_FMNMX3) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(e));
        } else if constexpr (KIND == K_MIX) { // per 2 cells: 2 max, 2 add, 1 max3
#pragma unroll
            for (int i = 0; i < CHAINS; i += 2) {
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i + 1]) : "f"(e));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i + 1]) : "f"(b[i + 1]));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(b[0]) : "f"(a[i]), "f"(a[i + 1]));
            }
        } else if constexpr (KIND == K_FADD2) {
#pragma unroll
            for (int i = 0; i < CHAINS; i += 2)
                asm volatile("{ .reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x; }"
                             : "+f"(a[i]), "+f"(a[i + 1])
                             : "f"(b[i]), "f"(b[i + 1]));
        } else if constexpr (KIND == K_MIX_FADD2) {
#pragma unroll
            for (int i = 0; i < CHAINS; i += 2) {
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(e));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i + 1]) : "f"(e));
                asm volatile("{ .reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x; }"
                             : "+f"(a[i]), "+f"(a[i + 1])
                             : "f"(b[i]), "f"(b[i + 1]));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(b[0]) : "f"(a[i]), "f"(a[i + 1]));
            }
        } else if constexpr (KIND == K_LDS128) {
#pragma unroll
            for (int i = 0; i < CHAINS; i += 4) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(static_cast<unsigned>(__cvta_generic_to_shared(lane4 + ((it + i) & 63) * 32))));
                a[i] += v.x;
                a[i + 1] += v.y;
                a[i + 2] += v.z;
                a[i + 3] += v.w;
            }
        } else if constexpr (KIND == K_MIX_LDS) { // the recurrence's real mix: LDS.128 + 4x(max, add) + 2x max3
#pragma unroll
            for (int i = 0; i < CHAINS; i += 4) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(static_cast<unsigned>(__cvta_generic_to_shared(lane4 + ((it + i) & 63) * 32))));
                a[i + 3] = v.w + fmaxf(a[i + 2], e);
                a[i + 2] = v.z + fmaxf(a[i + 1], e);
                a[i + 1] = v.y + fmaxf(a[i], e);
                a[i] = v.x + fmaxf(b[i], e);
                b[0] = fmaxf(fmaxf(b[0], a[i + 3]), a[i + 2]);
                b[1] = fmaxf(fmaxf(b[1], a[i + 1]), a[i]);
            }
        } else if constexpr (KIND == K_CREDUX) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                float r;
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(a[i]));
                a[i] = r + b[i];
            }
        } else if constexpr (KIND == K_SHFL) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1) + b[i];
        }
    }
    const long long t1 = clock64();
    float s = e;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- TMEM read throughput: tcgen05.ld 32 lanes x 32 bit x 16 columns per instruction ----
__global__ void __launch_bounds__(1024, 1) probe_tmem(float* out, long long* cycles, int wait_every) {
    __shared__ unsigned tmem_base_smem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            static_cast<unsigned>(__cvta_generic_to_shared(&tmem_base_smem))));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned base = tmem_base_smem + ((static_cast<unsigned>(warp & 3) * 32u) << 16);
    float acc = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        unsigned r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(base + ((it * 16) & 255)));
        if ((it % wait_every) == wait_every - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(r[0] & 0x3f800000u);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_smem));
}

template <int KIND> double run(int threads, float* out, long long* d_cycles, int sms, double lane_ops_per_iter_per_thread) {
    probe<KIND><<<sms, threads, 8192 * sizeof(float)>>>(out, d_cycles, 1.0f);
    CHECK(cudaDeviceSynchronize());
    probe<KIND><<<sms, threads, 8192 * sizeof(float)>>>(out, d_cycles, 1.0f);
    CHECK(cudaDeviceSynchronize());
    std::vector<long long> cyc(sms);
    CHECK(cudaMemcpy(cyc.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long worst = 0;
    for (auto c : cyc) worst = c > worst ? c : worst;
    return lane_ops_per_iter_per_thread * ITERS * threads / static_cast<double>(worst);
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float* out;
    long long* d_cycles;
    CHECK(cudaMalloc(&out, sizeof(float) * sms * 1024));
    CHECK(cudaMalloc(&d_cycles, sizeof(long long) * sms));
    printf("{\"device\": \"%s\", \"sms\": %d, \"unit\": \"lane-ops (or bytes) per clock per SM\"", prop.name, sms);
    for (int threads : {256, 512, 1024}) {
        printf(",\n \"threads_%d\": {", threads);
        printf("\"fadd\": %.1f", run<K_FADD>(threads, out, d_cycles, sms, CHAINS));
        printf(", \"fmnmx\": %.1f", run<K_FMNMX>(threads, out, d_cycles, sms, CHAINS));
        printf(", \"fmnmx3_instr\": %.1f", run<K_FMNMX3>(threads, out, d_cycles, sms, CHAINS));
        printf(", \"mix_cells\": %.2f", run<K_MIX>(threads, out, d_cycles, sms, CHAINS)); // 1 cell per chain per iter
        printf(", \"fadd2_lane_ops\": %.1f", run<K_FADD2>(threads, out, d_cycles, sms, CHAINS));
        printf(", \"mix_fadd2_cells\": %.2f", run<K_MIX_FADD2>(threads, out, d_cycles, sms, CHAINS));
        printf(", \"lds128_bytes\": %.1f", run<K_LDS128>(threads, out, d_cycles, sms, CHAINS * 4.0));
        printf(", \"mix_with_lds_cells\": %.2f", run<K_MIX_LDS>(threads, out, d_cycles, sms, CHAINS));
        printf(", \"credux_instr_lanes\": %.1f", run<K_CREDUX>(threads, out, d_cycles, sms, CHAINS));
        printf(", \"shfl_lanes\": %.1f", run<K_SHFL>(threads, out, d_cycles, sms, CHAINS));
        for (int wait_every : {1, 4}) {
            probe_tmem<<<sms, threads>>>(out, d_cycles, wait_every);
            CHECK(cudaDeviceSynchronize());
            probe_tmem<<<sms, threads>>>(out, d_cycles, wait_every);
            CHECK(cudaDeviceSynchronize());
            std::vector<long long> cyc(sms);
            CHECK(cudaMemcpy(cyc.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost));
            long long worst = 0;
            for (auto c : cyc) worst = c > worst ? c : worst;
            printf(", \"tmem_ld_x16_bytes_wait%d\": %.1f", wait_every, 64.0 * ITERS * threads / static_cast<double>(worst));
        }
        printf("}");
    }
    printf("\n}\n");
    return 0;
}
