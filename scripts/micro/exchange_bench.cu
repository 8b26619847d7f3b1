// Micro-benchmark of the barrier + reduction exchange of the persistent Krylov kernel (krylov_kernels.cuh, reduce()):
// n CTAs (one per SM, cooperative launch), each phase writes a few fields and exchanges N sums.  Prints cycles per exchange
// for the variants tried.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exchange_bench exchange_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <type_traits>

constexpr int T = 256, W = T / 32, NACC = 10, MAXG = 160;

__device__ __forceinline__ void st_release_stamped(double2* p, double v, unsigned long long s) {
    asm volatile("st.release.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(s) : "memory");
}
__device__ __forceinline__ void st_relaxed_stamped(double2* p, double v, unsigned long long s) {
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(s) : "memory");
}
__device__ __forceinline__ bool ld_relaxed_stamped(const double2* p, unsigned long long s, double& v) {
    long long b; unsigned long long g;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(b), "=l"(g) : "l"(p) : "memory");
    v = __longlong_as_double(b);
    return g == s;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// VARIANT 0: as in the kernel.  1: no field writes before (fence has nothing to drain).  2: one fence + relaxed stores.
// 3: polling by warp 0 only.  4: no trailing __threadfence (timing only; not a valid acquire).
// 5: record of N plain doubles + ONE stamp written with a release store; the stamp is polled, the record read once.
// 6: the first CTA collects (as 5), adds and publishes one copy of the totals per follower (record + stamp).
template <int N, int VARIANT>
__global__ void __launch_bounds__(T) bench(double2* partials, double2* field, int n_field, int reps, long long* cycles, double* out) {
    __shared__ double sh_part[W][NACC], sh_red[NACC], sh_all[MAXG][NACC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = gridDim.x;
    unsigned long long stamp = 1ull << 32;
    int parity = 0;
    double keep = 0.0;
    long long t0 = clock64(), spent = 0;
    for (int r = 0; r < reps; ++r) {
        if (VARIANT != 1) {   // the phase's output: a few 16-byte stores per thread
            for (int f = 0; f < 4; ++f) field[((size_t)f * n + blockIdx.x) * T + tid] = make_double2(r, f + keep);
        }
        double acc[N];
        for (int k = 0; k < N; ++k) acc[k] = tid * 1e-3 + k + r;
        const long long ta = clock64();
        for (int k = 0; k < N; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) for (int k = 0; k < N; ++k) sh_part[warp][k] = acc[k];
        __syncthreads();
        stamp += 1;
        double2* mine = partials + ((size_t)parity * n + blockIdx.x) * NACC;
        if (VARIANT == 2) {
            if (tid < N) {
                double s = 0.0;
                for (int w = 0; w < W; ++w) s += sh_part[w][tid];
                __threadfence();
                st_relaxed_stamped(mine + tid, s, stamp);
            }
        } else if (VARIANT < 5 && tid < N) {
            double s = 0.0;
            for (int w = 0; w < W; ++w) s += sh_part[w][tid];
            st_release_stamped(mine + tid, s, stamp);
        }
        const double2* base = partials + (size_t)parity * n * NACC;
        if (VARIANT == 5 || VARIANT == 6) {
            // record layout per CTA: NACC doubles + stamp, padded to 128 B (16 doubles)
            double* rec = reinterpret_cast<double*>(partials) + ((size_t)parity * 2 * n + blockIdx.x) * 16;
            unsigned long long* flag = reinterpret_cast<unsigned long long*>(rec + 15);
            __syncthreads();   // (the stamped stores above are part of the other variants; harmless here)
            if (tid < N) { double s2 = 0.0; for (int w = 0; w < W; ++w) s2 += sh_part[w][tid]; rec[tid] = s2; }
            __syncthreads();
            if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(stamp) : "memory");
            const bool collect = VARIANT == 5 || blockIdx.x == 0;
            if (collect) {
                const double* all = reinterpret_cast<const double*>(partials) + (size_t)parity * 2 * n * 16;
                for (int c = tid; c < n; c += T) {
                    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(all + (size_t)c * 16 + 15);
                    unsigned long long g;
                    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(g) : "l"(f) : "memory"); } while (g != stamp);
                    for (int k = 0; k < N; k += 2) {
                        const double2 v = __ldcg(reinterpret_cast<const double2*>(all + (size_t)c * 16 + k));
                        sh_all[c][k] = v.x; if (k + 1 < N) sh_all[c][k + 1] = v.y;
                    }
                }
                __syncthreads();
                if (warp == 0) {
                    double s[N];
                    for (int k = 0; k < N; ++k) s[k] = 0.0;
                    for (int c = lane; c < n; c += 32) for (int k = 0; k < N; ++k) s[k] += sh_all[c][k];
                    for (int k = 0; k < N; ++k) s[k] = warp_sum(s[k]);
                    if (lane == 0) for (int k = 0; k < N; ++k) sh_red[k] = s[k];
                }
                __syncthreads();
            }
            if (VARIANT == 6) {
                double* tot = reinterpret_cast<double*>(partials) + ((size_t)parity * 2 * n + n) * 16;   // second half: totals, one record per CTA
                if (blockIdx.x == 0) {
                    for (int e = tid; e < (n - 1) * N; e += T) { const int c = 1 + e / N, k = e - (c - 1) * N; tot[(size_t)c * 16 + k] = sh_red[k]; }
                    __syncthreads();
                    __threadfence();
                    for (int c = 1 + tid; c < n; c += T)
                        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(tot + (size_t)c * 16 + 15)), "l"(stamp) : "memory");
                } else {
                    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(tot + (size_t)blockIdx.x * 16 + 15);
                    if (tid == 0) { unsigned long long g; do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(g) : "l"(f) : "memory"); } while (g != stamp); }
                    __syncthreads();
                    if (tid < N) sh_red[tid] = __ldcg(tot + (size_t)blockIdx.x * 16 + tid);
                    __syncthreads();
                }
            }
            keep += sh_red[0] * 1e-30;
            parity ^= 1;
            spent += clock64() - ta;
            continue;
        }
        if (VARIANT == 3) {
            if (warp == 0) {
                for (int c = lane; c < n; c += 32) {
                    double v[N]; bool ok;
                    do { ok = true; for (int k = 0; k < N; ++k) ok &= ld_relaxed_stamped(base + (size_t)c * NACC + k, stamp, v[k]); } while (!ok);
                    for (int k = 0; k < N; ++k) sh_all[c][k] = v[k];
                }
                __threadfence();
            }
        } else {
            for (int c = tid; c < n; c += T) {
                double v[N]; bool ok;
                do { ok = true; for (int k = 0; k < N; ++k) ok &= ld_relaxed_stamped(base + (size_t)c * NACC + k, stamp, v[k]); } while (!ok);
                for (int k = 0; k < N; ++k) sh_all[c][k] = v[k];
            }
            if (VARIANT != 4) __threadfence();
        }
        __syncthreads();
        if (warp == 0) {
            double s[N];
            for (int k = 0; k < N; ++k) s[k] = 0.0;
            for (int c = lane; c < n; c += 32) for (int k = 0; k < N; ++k) s[k] += sh_all[c][k];
            for (int k = 0; k < N; ++k) s[k] = warp_sum(s[k]);
            if (lane == 0) for (int k = 0; k < N; ++k) sh_red[k] = s[k];
        }
        __syncthreads();
        keep += sh_red[0] * 1e-30;
        parity ^= 1;
        spent += clock64() - ta;
    }
    if (tid == 0) { cycles[blockIdx.x * 2] = spent; cycles[blockIdx.x * 2 + 1] = clock64() - t0; out[blockIdx.x] = keep; }
}

template <int N, int V>
void run(const char* name, int n, int reps) {
    double2 *partials, *field; long long* cycles; double* out;
    cudaMalloc(&partials, sizeof(double) * 16 * 4 * n + sizeof(double2) * 2 * n * NACC); cudaMemset(partials, 0, sizeof(double) * 16 * 4 * n + sizeof(double2) * 2 * n * NACC);
    cudaMalloc(&field, sizeof(double2) * 4 * n * T); cudaMalloc(&cycles, sizeof(long long) * 2 * n); cudaMalloc(&out, sizeof(double) * n);
    int nf = 4 * n * T;
    void* args[] = {&partials, &field, &nf, &reps, &cycles, &out};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchCooperativeKernel((const void*)bench<N, V>, dim3(n), dim3(T), args, 0, 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { std::printf("%s: launch failed\n", name); return; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)std::malloc(sizeof(long long) * 2 * n); cudaMemcpy(h, cycles, sizeof(long long) * 2 * n, cudaMemcpyDeviceToHost);
    double mean = 0; for (int c = 0; c < n; ++c) mean += double(h[c * 2]) / n;
    std::printf("%-58s n=%3d N=%2d: %7.0f cycles per exchange (mean over CTAs), %6.2f us per repetition\n", name, n, N, mean / reps, ms * 1e3 / reps);
    cudaFree(partials); cudaFree(field); cudaFree(cycles); cudaFree(out); std::free(h);
}

int main() {
    const int reps = 2000;
    for (int n : {8, 99, 148}) {
        run<2, 0>("as in the kernel", n, reps);
        run<10, 0>("as in the kernel", n, reps);
        run<10, 1>("nothing written before the exchange", n, reps);
        run<10, 2>("one fence + relaxed stamped stores", n, reps);
        run<10, 3>("polled by one warp", n, reps);
        run<10, 4>("no fence after polling (timing only)", n, reps);
        run<2, 5>("record + one stamp, all-to-all", n, reps);
        run<8, 5>("record + one stamp, all-to-all", n, reps);
        run<10, 5>("record + one stamp, all-to-all", n, reps);
        run<2, 6>("record + one stamp, first CTA adds and publishes", n, reps);
        run<8, 6>("record + one stamp, first CTA adds and publishes", n, reps);
    }
    return 0;
}
