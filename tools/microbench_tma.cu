// microbench_tma.cu -- how fast can a persistent grid (2 CTAs x 256 threads per SM, 100 KB of shared
// memory each, like pk_fwd_kernel) push tiles of S[b][f][p] (c2 shape: 257 rows x 32 slices x 8 B per
// tile, row pitch ldp) to HBM, by store mechanism:
//   tma16  : 16 boxes [32 slices][16 rows] + Nyquist box per tile (one per last-pass group)
//   tma1   : 1 box [32][16][16] + Nyquist box per tile
//   stg2   : STG.64, half-warp per row (two 128-byte runs per instruction)   (old plain-store path)
//   stg1   : STG.64, one row per warp instruction (256-byte run)
//   stg128 : STG.128, 16 lanes per row
//   bulk   : cp.async.bulk shared->global, one 256-byte row per instruction
// Output: GB/s of algorithmic bytes (257 * P * 8 * B).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    return (EncodeFn)p;
}
// dims {P, d: 16 (+1), g: 16, B}; row f = g + 16 d
static CUtensorMap make_map(void *base, int P, int ldp, int F, int B, int box_d, int box_g) {
    CUtensorMap m;
    const cuuint64_t dims[4] = {(cuuint64_t)P, 17, 16, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)16 * ldp * 8, (cuuint64_t)ldp * 8, (cuuint64_t)F * ldp * 8};
    const cuuint32_t box[4] = {32, (cuuint32_t)box_d, (cuuint32_t)box_g, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

__device__ __forceinline__ void tma_store(const CUtensorMap *m, unsigned src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

enum { TMA16 = 0, TMA1, STG2, STG1, STG128, BULK, LD_TMA, LD_LDG, LD_LDG2 };
__device__ __forceinline__ void tma_load(const CUtensorMap *m, unsigned dst, unsigned bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ int g_ld_off = 0;       // column offset of the TMA loads (alignment probe)
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float *S, int P, int ldp, int F, int ptiles, long total,
                                             const __grid_constant__ CUtensorMap m16, const __grid_constant__ CUtensorMap m1) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *w = reinterpret_cast<float *>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 257 * 64; i += 256) w[i] = i * 0.25f + blockIdx.x;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const unsigned wb = (unsigned)__cvta_generic_to_shared(w);
    __shared__ __align__(8) unsigned long long bar_mem;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bar_mem);
    if (MODE == LD_TMA && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    unsigned phase = 0;
    for (long t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = (int)(t / ptiles), pt = (int)(t % ptiles);
        const int p0 = pt * 32 + (MODE == LD_TMA ? g_ld_off : 0);
        if (MODE == TMA16) {
            if (lane == 0) {
                tma_store(&m16, wb + (2 * warp) * 4096, p0, 0, 2 * warp, b);
                tma_store(&m16, wb + (2 * warp + 1) * 4096, p0, 0, 2 * warp + 1, b);
                if (warp == 0) tma_store(&m16, wb + 65536, p0, 16, 0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else if (MODE == TMA1) {
            if (tid == 0) {
                tma_store(&m1, wb, p0, 0, 0, b);
                tma_store(&m16, wb + 65536, p0, 16, 0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else if (MODE == BULK) {
            // staged row r holds bin f = (r % 16) * 16 + r / 16 ... any bijection will do for timing: use f = r
            for (int r = tid; r < 257; r += 256) {
                if (p0 + 32 <= P) {
                    char *dst = reinterpret_cast<char *>(S) + (((size_t)b * F + r) * ldp + p0) * 8;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" ::"l"(dst), "r"(wb + r * 256) : "memory");
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else if (MODE == LD_TMA) {
            // 17 boxes [16 rows][32 slices] -> shared memory, completion on one mbarrier; then every
            // lane reads its slice of the 33 rows of its warp (what the paired inverse pass 1 needs)
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(17 * 4096) : "memory");   // OOB rows of a box are zero-filled and counted
                for (int g = 0; g < 16; ++g) tma_load(&m16, wb + g * 4096, bar, p0, 0, g, b);
                tma_load(&m16, wb + 65536, bar, p0, 16, 0, b);
            }
            asm volatile(
                "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                ::"r"(bar), "r"(phase) : "memory");
            phase ^= 1;
            for (int r = warp * 32; r < warp * 32 + 32; ++r) {
                const float2 v = *reinterpret_cast<const float2 *>(w + r * 64 + 2 * lane);
                acc += v.x + v.y;
            }
        } else if (MODE == LD_LDG || MODE == LD_LDG2) {
            // LD_LDG: one 256-byte row run per warp instruction (paired pass 1); LD_LDG2: two 128-byte
            // runs of two rows per instruction (the two-slices-per-lane pass 1)
            if (MODE == LD_LDG) {
                for (int r = warp * 32; r < warp * 32 + 32 + (warp == 0); ++r) {
                    const int rr = r < 257 ? r : 256;
                    const char *row = reinterpret_cast<const char *>(S) + (((size_t)b * F + rr) * ldp + p0) * 8;
                    if (p0 + lane < P) { const float2 v = __ldg(reinterpret_cast<const float2 *>(row + 8 * lane)); acc += v.x + v.y; }
                }
            } else {
                const int l = lane & 15, sub = lane >> 4;
                for (int r = warp * 2 + sub; r < 257; r += 16) {
                    const char *row = reinterpret_cast<const char *>(S) + (((size_t)b * F + r) * ldp + p0) * 8;
                    if (p0 + l < P) { const float2 v = __ldg(reinterpret_cast<const float2 *>(row + 8 * l)); acc += v.x + v.y; }
                    if (p0 + l + 16 < P) { const float2 v = __ldg(reinterpret_cast<const float2 *>(row + 8 * (l + 16))); acc += v.x + v.y; }
                }
            }
        } else if (MODE == STG2) {
            const int l = lane & 15, sub = lane >> 4;
            for (int r = warp * 2 + sub; r < 257; r += 16) {
                char *row = reinterpret_cast<char *>(S) + (((size_t)b * F + r) * ldp + p0) * 8;
                const float2 v0 = *reinterpret_cast<const float2 *>(w + r * 64 + 2 * l);
                const float2 v1 = *reinterpret_cast<const float2 *>(w + r * 64 + 2 * (l + 16));
                if (p0 + l < P) *reinterpret_cast<float2 *>(row + 8 * l) = v0;
                if (p0 + l + 16 < P) *reinterpret_cast<float2 *>(row + 8 * (l + 16)) = v1;
            }
        } else if (MODE == STG1) {
            for (int r = warp; r < 257; r += 8) {
                char *row = reinterpret_cast<char *>(S) + (((size_t)b * F + r) * ldp + p0) * 8;
                const float2 v0 = *reinterpret_cast<const float2 *>(w + r * 64 + 2 * lane);
                if (p0 + lane < P) *reinterpret_cast<float2 *>(row + 8 * lane) = v0;
            }
        } else if (MODE == STG128) {
            const int l = lane & 15, sub = lane >> 4;
            for (int r = warp * 2 + sub; r < 257; r += 16) {
                char *row = reinterpret_cast<char *>(S) + (((size_t)b * F + r) * ldp + p0) * 8;
                const float4 v0 = *reinterpret_cast<const float4 *>(w + r * 64 + 4 * l);
                if (p0 + 2 * l + 1 < P) *reinterpret_cast<float4 *>(row + 16 * l) = v0;
            }
        }
        __syncthreads();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 12345.678f) S[0] = acc;
}

template <int MODE>
static void run(const char *name, float *S, int P, int ldp, int F, int B, int nsm) {
    const int ptiles = (P + 31) / 32;
    const long total = (long)ptiles * B;
    CUtensorMap m16 = make_map(S, P, ldp, F, B, 16, 1), m1 = make_map(S, P, ldp, F, B, 16, 16);
    const int smem = 100 * 1024;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(a));
        k<MODE><<<2 * nsm, 256, smem>>>(S, P, ldp, F, ptiles, total, m16, m1);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-8s ldp=%d  %.3f ms  %.1f GB/s\n", name, ldp, best, (double)B * F * P * 8 / best / 1e6);
}

int main() {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    const int B = 1024, F = 257, P = 1253;
    float *S; CK(cudaMalloc(&S, (size_t)B * F * 1280 * 8 + 65536));
    for (int off : {0, 2, 1, -3}) {      // does the start column of a tensor load have to be 16-byte aligned?
        CK(cudaMemcpyToSymbol(g_ld_off, &off, sizeof(int)));
        printf("ld_tma column offset %d: ", off);
        run<LD_TMA>("ld_tma", S, P, 1256, F, B, nsm);
    }
    { int off = 0; CK(cudaMemcpyToSymbol(g_ld_off, &off, sizeof(int))); }
    for (int ldp : {1256, 1280}) {
        run<TMA16>("tma16", S, P, ldp, F, B, nsm);
        run<TMA1>("tma1", S, P, ldp, F, B, nsm);
        run<BULK>("bulk", S, P, ldp, F, B, nsm);
        run<STG2>("stg2", S, P, ldp, F, B, nsm);
        run<STG1>("stg1", S, P, ldp, F, B, nsm);
        run<STG128>("stg128", S, P, ldp, F, B, nsm);
        run<LD_TMA>("ld_tma", S, P, ldp, F, B, nsm);
        run<LD_LDG>("ld_ldg1", S, P, ldp, F, B, nsm);
        run<LD_LDG2>("ld_ldg2", S, P, ldp, F, B, nsm);
    }
    return 0;
}
