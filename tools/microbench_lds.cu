// microbench_lds.cu -- what a shared-memory / shuffle / constant access costs on the LSU data
// pipe of a B200 SM, for the access shapes the slice-lane kernels use:
//   * broadcast table reads (all 32 lanes one address / one address per half-warp) as
//     LDS.32 / LDS.64 / LDS.128
//   * conflict-free data reads LDS.32 / LDS.64 / LDS.128 (distinct addresses)
//   * SHFL
//   * LDC with a per-warp / per-half-warp index
// One CTA per SM, 16 warps, every warp runs `iters` x 16 independent accesses; reported:
// SM cycles per warp-instruction (= 1 / throughput), from clock64 around the loop.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__constant__ float4 ctab[1024];

enum { K_LDS32 = 0, K_LDS64, K_LDS128, K_SHFL, K_LDC128 };

// pattern: 0 = all lanes same address, 1 = one address per half-warp, 2 = distinct per lane
// (conflict-free), 3 = one address per quarter-warp
template <int KIND, int PAT>
__global__ void __launch_bounds__(512, 1) k(float *out, long long *cyc, int iters) {
    extern __shared__ float4 sm4[];
    float *sm = reinterpret_cast<float *>(sm4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = i * 0.5f;
    __syncthreads();
    int sel;
    if (PAT == 0) sel = 0;
    else if (PAT == 1) sel = lane >> 4;
    else if (PAT == 3) sel = lane >> 3;
    else sel = lane;
    const int esz = KIND == K_LDS32 ? 1 : KIND == K_LDS64 ? 2 : 4;       // floats
    // per-instruction address offsets: 16 different rows (keeps the loads independent)
    int base = (warp * 37 + sel) * esz;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int a = (base + u * 68 * esz) & 8191 & ~(esz - 1);
            if (KIND == K_LDS32) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + a)));
                acc0 += v;
            } else if (KIND == K_LDS64) {
                float v0, v1;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0), "=f"(v1) : "r"((unsigned)__cvta_generic_to_shared(sm + a)));
                acc0 += v0; acc1 += v1;
            } else if (KIND == K_LDS128) {
                float v0, v1, v2, v3;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"((unsigned)__cvta_generic_to_shared(sm + a)));
                acc0 += v0; acc1 += v1; acc2 += v2; acc3 += v3;
            } else if (KIND == K_SHFL) {
                acc0 += __shfl_xor_sync(0xffffffffu, acc1 + u, 16);
            } else if (KIND == K_LDC128) {
                const float4 v = ctab[(a >> 2) & 1023];
                acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
            }
        }
        base += 4 * esz;
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    const float s = acc0 + acc1 + acc2 + acc3;
    if (s == 12345.678f) out[0] = s;
}

template <int KIND, int PAT>
static void run(const char *name, float *out, long long *cyc, int nsm) {
    const int iters = 2000, threads = 512;
    CK(cudaFuncSetAttribute(k<KIND, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    k<KIND, PAT><<<nsm, threads, 65536>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    k<KIND, PAT><<<nsm, threads, 65536>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < nsm; ++i) avg += (double)h[i];
    avg /= nsm;
    const double ninst = (double)iters * 16 * (threads / 32);
    printf("%-36s %.3f cycles per warp-instruction per SM\n", name, avg / ninst);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    float *out; long long *cyc;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, sizeof(long long) * 256));
    static float4 h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = make_float4(i, i + 1, i + 2, i + 3);
    CK(cudaMemcpyToSymbol(ctab, h, sizeof(h)));
    run<K_LDS32, 0>("LDS.32  all lanes one address", out, cyc, nsm);
    run<K_LDS32, 1>("LDS.32  one address per half-warp", out, cyc, nsm);
    run<K_LDS32, 2>("LDS.32  distinct (128 B)", out, cyc, nsm);
    run<K_LDS64, 0>("LDS.64  all lanes one address", out, cyc, nsm);
    run<K_LDS64, 1>("LDS.64  one address per half-warp", out, cyc, nsm);
    run<K_LDS64, 3>("LDS.64  one address per quarter-warp", out, cyc, nsm);
    run<K_LDS64, 2>("LDS.64  distinct (256 B)", out, cyc, nsm);
    run<K_LDS128, 0>("LDS.128 all lanes one address", out, cyc, nsm);
    run<K_LDS128, 1>("LDS.128 one address per half-warp", out, cyc, nsm);
    run<K_LDS128, 3>("LDS.128 one address per quarter-warp", out, cyc, nsm);
    run<K_LDS128, 2>("LDS.128 distinct (512 B)", out, cyc, nsm);
    run<K_SHFL, 0>("SHFL.BFLY", out, cyc, nsm);
    run<K_LDC128, 0>("LDC.128 all lanes one index", out, cyc, nsm);
    run<K_LDC128, 1>("LDC.128 one index per half-warp", out, cyc, nsm);
    return 0;
}
