// microbench.cu -- hardware questions behind the kernel design (run on the B200 via gpurun):
//   1. FP32 FMA issue rate: scalar FFMA vs packed fma.rn.f32x2 (FFMA2); FP64 DFMA
//   2. HBM write / read bandwidth for the reference's [f_pts][P] layout when a CTA owns a tile
//      of T consecutive slices (row runs of T*8 bytes at 8-byte alignment), vs a flat copy
//   3. shared-memory bandwidth (LDS.64 / LDS.128)
// Output: one line per measurement, "name value unit".
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}

// ---- 1. FMA rates ---------------------------------------------------------------------------
__global__ void ffma_scalar(float *out, int iters, float a, float b) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 12345.678f) out[0] = s;
}
__global__ void ffma_packed(float *out, int iters, float a, float b) {
    unsigned long long r[8];
    unsigned long long pa, pb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(r[i]) : "f"(x)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(pa), "l"(pb));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i])); s += lo + hi; }
    if (s == 12345.678f) out[0] = s;
}
__global__ void fadd_packed(float *out, int iters, float a) {
    unsigned long long r[8], pa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(r[i]) : "f"(x)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(pa));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i])); s += lo + hi; }
    if (s == 12345.678f) out[0] = s;
}
__global__ void fadd_scalar(float *out, int iters, float a) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = r[i] + a;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 12345.678f) out[0] = s;
}
__global__ void dfma_scalar(double *out, int iters, double a, double b) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 0.001 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    if (s == 12345.678) out[0] = s;
}

// ---- 2. [F][P] tile stores / loads ------------------------------------------------------------
// CTA = (signal b, tile of T slices); writes rows k = 0..F-1, each a run of T complex<float>.
// LPR lanes per row-run, each lane writes V floats (V = 2 or 4).
template <int T, int V>
__global__ void tile_store(float *S, int F, int P, int ldp, int ptiles) {
    constexpr int LPR = T * 2 / V;              // lanes per row run
    constexpr int RPW = 32 / LPR;               // rows per warp instruction
    const int b = blockIdx.x / ptiles, pt = blockIdx.x % ptiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int rsub = lane / LPR, l = lane % LPR;
    const int p = pt * T + l * (V / 2);
    float val = (float)(threadIdx.x + blockIdx.x);
    for (int k = warp * RPW + rsub; k < F; k += nw * RPW) {
        if (p < P) {
            float *dst = S + (((size_t)b * F + k) * ldp + p) * 2;
            if (V == 2) { *(float2 *)dst = make_float2(val, val + k); }
            else {
                if (p + 1 < P) {
                    // 16-byte store only if aligned, else two 8-byte stores
                    if ((((size_t)dst) & 15) == 0) *(float4 *)dst = make_float4(val, val + k, val, val - k);
                    else { *(float2 *)dst = make_float2(val, val + k); *(float2 *)(dst + 2) = make_float2(val, val - k); }
                } else *(float2 *)dst = make_float2(val, val + k);
            }
        }
    }
}
// the packed kernels' store shapes for a tile of 32 slices: half-warp (16 lanes) per row, 2 rows per
// instruction.  PAIR = 0: lane l holds slices (l, l+16): two 8-byte stores, 128-byte runs each;
// PAIR = 1: lane l holds slices (2l, 2l+1): one 16-byte store (two 8-byte ones if misaligned), 256-byte runs.
template <int PAIR>
__global__ void tile_store_hw(float *S, int F, int P, int ldp, int ptiles) {
    const int b = blockIdx.x / ptiles, pt = blockIdx.x % ptiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int l = lane & 15, sub = lane >> 4;
    float val = (float)(threadIdx.x + blockIdx.x);
    for (int k = warp * 2 + sub; k < F; k += nw * 2) {
        float *row = S + (((size_t)b * F + k) * ldp + pt * 32) * 2;
        if (PAIR == 0) {
            if (pt * 32 + l < P) *(float2 *)(row + 2 * l) = make_float2(val, val + k);
            if (pt * 32 + l + 16 < P) *(float2 *)(row + 2 * (l + 16)) = make_float2(val, val - k);
        } else {
            float *dst = row + 4 * l;
            if (pt * 32 + 2 * l + 1 < P) {
                if ((((size_t)dst) & 15) == 0) *(float4 *)dst = make_float4(val, val + k, val, val - k);
                else { *(float2 *)dst = make_float2(val, val + k); *(float2 *)(dst + 2) = make_float2(val, val - k); }
            } else if (pt * 32 + 2 * l < P) *(float2 *)dst = make_float2(val, val + k);
        }
    }
}
template <int T>
__global__ void tile_load(const float *S, float *sink, int F, int P, int ldp, int ptiles) {
    constexpr int LPR = T;
    constexpr int RPW = 32 / LPR > 0 ? 32 / LPR : 1;
    const int b = blockIdx.x / ptiles, pt = blockIdx.x % ptiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int rsub = lane / LPR, l = lane % LPR;
    const int p = pt * T + l;
    float acc = 0;
    for (int k = warp * RPW + rsub; k < F; k += nw * RPW) {
        if (p < P) {
            const float2 v = *(const float2 *)(S + (((size_t)b * F + k) * ldp + p) * 2);
            acc += v.x + v.y;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}
__global__ void flat_copy(const float4 *a, float4 *b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void flat_write(float4 *b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_float4(1, 2, 3, (float)i);
}
__global__ void flat_read(const float4 *a, float *sink, size_t n) {
    float acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = a[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) sink[0] = acc;
}

// ---- 3. shared memory bandwidth -----------------------------------------------------------------
template <int V>
__global__ void smem_bw(float *out, int iters) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float acc = 0;
    int idx = threadIdx.x * V;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int a = (idx + u * 1024) & 8191;
            if (V == 2) { float2 v = *(float2 *)&sm[a]; acc += v.x + v.y; }
            else { float4 v = *(float4 *)&sm[a]; acc += v.x + v.y + v.z + v.w; }
        }
        idx = (idx + 64) & 8191;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int T, int V>
static void run_store(const char *name, float *S, int B, int F, int P, int ldp) {
    const int ptiles = (P + T - 1) / T;
    float ms = time_ms([&] { tile_store<T, V><<<B * ptiles, 256>>>(S, F, P, ldp, ptiles); });
    printf("%s T=%d V=%d ldp=%d  %.1f GB/s\n", name, T, V, ldp, (double)B * F * P * 8 / ms / 1e6);
}
template <int T>
static void run_load(const char *name, float *S, float *sink, int B, int F, int P, int ldp) {
    const int ptiles = (P + T - 1) / T;
    float ms = time_ms([&] { tile_load<T><<<B * ptiles, 256>>>(S, sink, F, P, ldp, ptiles); });
    printf("%s T=%d ldp=%d  %.1f GB/s\n", name, T, ldp, (double)B * F * P * 8 / ms / 1e6);
}

int main() {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s SMs %d clock %d kHz L2 %d MB smem/SM %zu\n", pr.name, pr.multiProcessorCount, pr.clockRate, pr.l2CacheSize >> 20, pr.sharedMemPerMultiprocessor);
    float *out; CK(cudaMalloc(&out, 1024));
    const int SMS = pr.multiProcessorCount;
    {
        const int iters = 4096, blocks = SMS * 8, thr = 256;
        float ms = time_ms([&] { ffma_scalar<<<blocks, thr>>>(out, iters, 1.0001f, 0.5f); });
        printf("ffma_scalar %.1f TFLOP/s\n", 2.0 * blocks * thr * 16.0 * iters / ms / 1e9);
        ms = time_ms([&] { ffma_packed<<<blocks, thr>>>(out, iters, 1.0001f, 0.5f); });
        printf("ffma_packed_f32x2 %.1f TFLOP/s\n", 2.0 * blocks * thr * 16.0 * iters / ms / 1e9);
        ms = time_ms([&] { fadd_scalar<<<blocks, thr>>>(out, iters, 0.5f); });
        printf("fadd_scalar %.1f Tadd/s\n", 1.0 * blocks * thr * 16.0 * iters / ms / 1e9);
        ms = time_ms([&] { fadd_packed<<<blocks, thr>>>(out, iters, 0.5f); });
        printf("fadd_packed_f32x2 %.1f Tadd/s\n", 1.0 * blocks * thr * 16.0 * iters / ms / 1e9);
        ms = time_ms([&] { dfma_scalar<<<blocks, thr>>>((double *)out, iters, 1.0001, 0.5); });
        printf("dfma %.1f TFLOP/s\n", 2.0 * blocks * thr * 8.0 * iters / ms / 1e9);
    }
    {
        const int iters = 2048, blocks = SMS * 4, thr = 512;
        CK(cudaFuncSetAttribute(smem_bw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        CK(cudaFuncSetAttribute(smem_bw<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        float ms = time_ms([&] { smem_bw<2><<<blocks, thr, 32768>>>(out, iters); });
        printf("smem LDS.64  %.1f B/clk/SM (at %d kHz nominal)\n", (double)blocks * thr * iters * 8 * 8 / (ms * 1e-3) / SMS / (pr.clockRate * 1e3), pr.clockRate);
        ms = time_ms([&] { smem_bw<4><<<blocks, thr, 32768>>>(out, iters); });
        printf("smem LDS.128 %.1f B/clk/SM\n", (double)blocks * thr * iters * 8 * 16 / (ms * 1e-3) / SMS / (pr.clockRate * 1e3));
    }
    {
        // c2-like: B x 257 x 1253 complex64 (B = 512 -> 1.32 GB)
        const int B = 512, F = 257, P = 1253;
        const size_t cap = (size_t)B * F * 1280 * 8;
        float *S; CK(cudaMalloc(&S, cap)); CK(cudaMemset(S, 0, cap));
        float *S2; CK(cudaMalloc(&S2, cap));
        const size_t n4 = cap / 16;
        float ms = time_ms([&] { flat_copy<<<SMS * 16, 512>>>((const float4 *)S, (float4 *)S2, n4); });
        printf("flat_copy %.1f GB/s (read+write)\n", 2.0 * cap / ms / 1e6);
        ms = time_ms([&] { flat_write<<<SMS * 16, 512>>>((float4 *)S2, n4); });
        printf("flat_write %.1f GB/s\n", 1.0 * cap / ms / 1e6);
        ms = time_ms([&] { flat_read<<<SMS * 16, 512>>>((const float4 *)S, out, n4); });
        printf("flat_read %.1f GB/s\n", 1.0 * cap / ms / 1e6);
        ms = time_ms([&] { CK(cudaMemcpyAsync(S2, S, cap, cudaMemcpyDeviceToDevice)); });
        printf("cudaMemcpy D2D %.1f GB/s (read+write)\n", 2.0 * cap / ms / 1e6);
        for (int ldp : {1253, 1256, 1280}) {
            run_store<8, 2>("tile_store", S, B, F, P, ldp);
            run_store<16, 2>("tile_store", S, B, F, P, ldp);
            run_store<32, 2>("tile_store", S, B, F, P, ldp);
            run_store<64, 4>("tile_store", S, B, F, P, ldp);
        }
        for (int ldp : {1253, 1254, 1256}) {
            const int ptiles = (P + 31) / 32;
            float ms = time_ms([&] { tile_store_hw<0><<<B * ptiles, 256>>>(S, F, P, ldp, ptiles); });
            printf("tile_store_hw pair(l,l+16) 2x8B ldp=%d  %.1f GB/s\n", ldp, (double)B * F * P * 8 / ms / 1e6);
            ms = time_ms([&] { tile_store_hw<1><<<B * ptiles, 256>>>(S, F, P, ldp, ptiles); });
            printf("tile_store_hw pair(2l,2l+1) 16B ldp=%d  %.1f GB/s\n", ldp, (double)B * F * P * 8 / ms / 1e6);
        }
        for (int ldp : {1253, 1256}) {
            run_load<8>("tile_load", S, out, B, F, P, ldp);
            run_load<16>("tile_load", S, out, B, F, P, ldp);
            run_load<32>("tile_load", S, out, B, F, P, ldp);
        }
        CK(cudaFree(S)); CK(cudaFree(S2));
    }
    return 0;
}
