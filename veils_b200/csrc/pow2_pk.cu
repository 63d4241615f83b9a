// pow2_pk.cu -- CUDA kernels and launchers of the packed f32 family (two slices per lane,
// FADD2/FMUL2/FFMA2).  Phase code: pow2_pk_impl.cuh.  Persistent CTAs walk the (signal, tile)
// list; plan tables live in shared memory and the signal tile is double-buffered with cp.async
// for mfft <= 512.
#include <cstdlib>
#include <cstring>

#include "pow2_pk_cfg.hpp"
#include "tma.cuh"
#include "launch_cache.cuh"

namespace veils {

namespace {

using namespace p2;

__device__ __forceinline__ void copy_table_f(float *dst, const float *src, int n, int tid, int nt) {
    for (int i = tid; i < n; i += nt) dst[i] = __ldg(src + i);
}

// co-resident CTAs the register budget is sized for: what the work buffer allows, capped so
// that every thread keeps 128 registers (512 threads per SM)
template <int LOGMC, int TF, int NT>
constexpr int pk_min_ctas() {
    constexpr int by_smem = 8 * (1 << LOGMC) * TF <= 40 * 1024 ? 4 : (8 * (1 << LOGMC) * TF <= 80 * 1024 ? 2 : 1);
    constexpr int by_regs = 512 / NT < 1 ? 1 : 512 / NT;
    return by_smem < by_regs ? by_smem : by_regs;
}

// STAGE (one-sided layouts): the last pass overwrites its own work rows with the finished bins
// -- every thread writes exactly the bytes it has read, so no barrier is needed -- and each
// half-warp hands its group's dense [RL rows][TF slices] box to the TMA engine (tma.cuh): the
// global stores cost the compute warps one instruction per group and drain asynchronously.
template <int LOGMC, int TF, int NSUB, bool ONES, bool SPEC, bool STAGE>
__global__ void __launch_bounds__((TF / 2) * NSUB, pk_min_ctas<LOGMC, TF, (TF / 2) * NSUB>())
pk_fwd_kernel(const __grid_constant__ FwdArgs<float> args, const PTab *__restrict__ ptab_g,
              const int64_t total, const int xin_elems, const int zero_off, const int stagger_ns,
              const int num_sms, const __grid_constant__ CUtensorMap smap,
              const __grid_constant__ CUtensorMap smap_tile) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using K = PkFwd<LOGMC, TF, NSUB>;
    const int tid = threadIdx.x;
    PkFwdCtx c;
    c.a = args;
    c.work = reinterpret_cast<float *>(smem);
    float *xin0 = c.work + (size_t)K::MC * 2 * TF + PK_NYQ_ROW;     // + staged Nyquist row
    float *xin1 = K::DBUF ? xin0 + xin_elems : xin0;
    if (K::STAB) {
        float *stw = xin0 + (K::DBUF ? 2 : 1) * xin_elems;     // xin_elems is a multiple of 4
        PTab *sptab = reinterpret_cast<PTab *>(stw + K::TWN_FWD);
        copy_table_f(stw, args.tw, K::TWN_FWD, tid, K::NT);
        copy_table_f(reinterpret_cast<float *>(sptab), reinterpret_cast<const float *>(ptab_g), 4 * K::MC, tid, K::NT);
        c.tw = stw;
        c.ptab = sptab;
    } else {
        c.tw = args.tw;
        c.ptab = ptab_g;
    }
    constexpr bool SPLIT1 = STAGE && K::L1 == NSUB;
    // 2-pass plans, complex output: the staged rows of all groups are contiguous (row g RL + d), so the
    // whole tile leaves as ONE box [G][RL][TF] (+ the Nyquist row), requested by one thread after the
    // barrier at the top of the next iteration -- a TMA instruction costs the issuing warp ~120 cycles,
    // 17 per tile spread over the warps' last passes cost ~9 % of the kernel
    constexpr bool ONEBOX = SPLIT1 && !SPEC;   // 3-pass plans: the same through a 5-D map (tma::make_s_map5)
    int64_t pb = 0, ppt = 0;               // previous tile (ONEBOX)
    const unsigned drain_bar = tma::smem_u32(c.work + (size_t)K::MC * 2 * TF + PK_NYQ_ROW - 4);
    if (SPLIT1 && tid == 0) {
        tma::mbar_init(drain_bar, K::NT / TF);
        tma::fence_mbar_init();
    }
    if (tid < 2) {                         // the zero pair behind the staged span
        xin0[zero_off + tid] = 0.f;
        xin1[zero_off + tid] = 0.f;
    }
    // (signal, tile) of work item w, advanced incrementally: w += gridDim.x
    // co-resident CTAs (blockIdx differing by the SM count) would otherwise run their
    // shared-memory-heavy and FP-heavy phases in lock step; a start offset interleaves them
    if (stagger_ns > 0 && (int)blockIdx.x >= num_sms) __nanosleep((unsigned)stagger_ns);
    int64_t w = blockIdx.x;
    int64_t cb = w / args.ptiles, cpt = w - cb * args.ptiles;
    const int64_t gb = gridDim.x / args.ptiles, gr = gridDim.x - gb * args.ptiles;
    if (w < total) {
        c.b = cb;
        c.pt = cpt;
        c.xin = xin0;
        K::fill(c, tid);
        cp_async_commit();
    }
    for (int it = 0; w < total; w += gridDim.x, ++it) {
        cp_async_wait_all();
        // STAGE: the previous tile's boxes must have left the work buffer before pass 1 overwrites
        // it.  One group per thread: that wait moves behind the tile reads + DFT of pass 1 (which
        // do not touch the work buffer) and is handed to the other threads through an mbarrier.
        if (STAGE && !SPLIT1) tma::wait_read_all();
        __syncthreads();
        if (ONEBOX && it > 0 && tid == 0) {
            const unsigned wbase = tma::smem_u32(c.work);
            if (K::NP == 2) tma::store_4d(&smap_tile, wbase, (int)(ppt * TF), 0, 0, (int)pb);
            else tma::store_5d(&smap_tile, wbase, (int)(ppt * TF), 0, 0, 0, (int)pb);
            tma::store_4d(&smap, wbase + K::MC * (TF * 8), (int)(ppt * TF), K::RL, 0, (int)pb);
            tma::commit_group();
        }
        float *cur = (K::DBUF && (it & 1)) ? xin1 : xin0;
        const int64_t wn = w + gridDim.x;
        int64_t nb = cb + gb, npt = cpt + gr;
        if (npt >= args.ptiles) { npt -= args.ptiles; ++nb; }
        if (K::DBUF && wn < total) {
            c.b = nb;
            c.pt = npt;
            c.xin = (it & 1) ? xin0 : xin1;
            K::fill(c, tid);
            cp_async_commit();
        }
        c.b = cb;
        c.pt = cpt;
        c.xin = cur;
#ifdef VEILS_ABLATE
        if (args.dbg & 4) {
        } else
#endif
        if (SPLIT1) {
            C2 v[K::R1];
            K::pass1_in(c, tid, tid / K::LN, v);
            if (tid % TF == 0) {
                tma::wait_read_all();          // this thread's boxes of the previous tile
                tma::mbar_arrive(drain_bar);
            }
            tma::mbar_wait(drain_bar, (unsigned)(it & 1));
            K::pass1_out(c, tid, tid / K::LN, v);
        } else {
            K::pass1(c, tid);
        }
        __syncthreads();
        if (!K::DBUF && wn < total) {
            PkFwdCtx cn = c;
            cn.b = nb;
            cn.pt = npt;
            K::fill(cn, tid);
            cp_async_commit();
        }
        if (K::NP == 3) {
            K::pass_mid(c, tid);
            __syncthreads();
        }
        if (STAGE) {
#pragma unroll 1
            for (int i = 0; i < K::WPT2; ++i) {
                C2 u[K::RL];
#ifdef VEILS_ABLATE
                if (args.dbg & 8) continue;
#endif
#ifdef VEILS_ABLATE
                if (!(args.dbg & 16)) {        // ablation 16: only the TMA stores remain of the last pass
#endif
                K::last_dft2(c, tid, i, u);
                __syncwarp();                  // the rows of a group pair belong to the threads of one warp:
                                               // all of them have read before anyone overwrites
                K::template last_emit2<SPEC>(c, tid, i, u);
#ifdef VEILS_ABLATE
                }
#endif
                tma::fence_async_smem();
                if (ONEBOX) continue;           // requested after the next barrier, by one thread
                __syncwarp();
                const int lam = tid % TF, wg = tid / TF + i * K::NW2;
#ifdef VEILS_ABLATE
                if (args.dbg & 1) continue;     // ablation: nothing leaves the SM
#endif
                if (lam == 0 && wg < K::G / 2) {
                    // one box [RL rows][TF slices] per group of the pair (+ the Nyquist row with pair 0).
                    // (1-D bulk copies, one per row, have the higher ceiling in a store-only kernel --
                    // 6.0 vs 5.0 TB/s, profiles/r1b_store_paths.txt -- but their per-lane issue loop
                    // costs more here than it gains.)
                    const unsigned wbase = tma::smem_u32(c.work);
                    const int p0 = (int)(cpt * TF);
                    const int g0 = K::group_of(wg, 0), g1 = K::group_of(wg, 1);
                    tma::store_4d(&smap, wbase + K::GEO::goff(g0) * (TF * 8), p0, 0, g0, (int)cb);
                    tma::store_4d(&smap, wbase + K::GEO::goff(g1) * (TF * 8), p0, 0, g1, (int)cb);
                    if (wg == 0)                // Nyquist row: rows past d = RL are clipped by the map
                        tma::store_4d(&smap, wbase + K::MC * (TF * 8), p0, K::RL, 0, (int)cb);
                    tma::commit_group();
                }
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < K::WPT; ++i) {
                C2 u[K::RL];
#ifdef VEILS_ABLATE
                if (args.dbg & 8) continue;
#endif
                K::last_dft(c, tid, i, u);
                K::template last_emit_v<ONES, SPEC, false>(c, tid, i, u, nullptr);
            }
        }
        pb = cb;
        ppt = cpt;
        cb = nb;
        cpt = npt;
    }
    if (ONEBOX) {
        __syncthreads();
        if (blockIdx.x < total && tid == 0) {  // the last tile of this CTA
            const unsigned wbase = tma::smem_u32(c.work);
            if (K::NP == 2) tma::store_4d(&smap_tile, wbase, (int)(ppt * TF), 0, 0, (int)pb);
            else tma::store_5d(&smap_tile, wbase, (int)(ppt * TF), 0, 0, 0, (int)pb);
            tma::store_4d(&smap, wbase + K::MC * (TF * 8), (int)(ppt * TF), K::RL, 0, (int)pb);
            tma::commit_group();
        }
    }
    if (STAGE) tma::wait_all();
}

// Request the S tile whose slice 0 is column `col0` of signal b, landing over the work buffer: ONE box
// [L1 groups][R1 rows][TF + 2 slices] (shared-memory row j R1 + a = bin j + L1 a) plus the Nyquist
// row.  (One box per group costs the issuing warp ~120 cycles per TMA instruction -- 17 of them held
// the whole CTA at the next barrier for 16 % of the kernel time, profiles/r1e_ncu_full_bench_c2.txt.)
// The start column is rounded down to even: 16-byte aligned.
template <typename K>
__device__ __forceinline__ void pk_land(const CUtensorMap &smap, const CUtensorMap &smap_nyq, const float *work,
                                        unsigned land_bar, int tid, int64_t b, int64_t col0) {
    if (tid != 0) return;
    tma::fence_async_smem();                   // earlier generic-proxy accesses to this buffer are ordered first
    const unsigned wbase = tma::smem_u32(work);
    const int col = (int)(col0 & ~(int64_t)1);
    tma::mbar_expect_tx(land_bar, (unsigned)K::LAND_BYTES);
    tma::load_4d(&smap, wbase, land_bar, col, 0, 0, (int)b);
    tma::load_4d(&smap_nyq, wbase + K::MC * (K::LTF * 8), land_bar, col, K::R1, 0, (int)b);
}

template <int LOGMC, int TF, int NSUB, bool ONES>
__global__ void __launch_bounds__((TF / 2) * NSUB, pk_min_ctas<LOGMC, TF, (TF / 2) * NSUB>())
pk_inv_kernel(const __grid_constant__ InvArgs<float> args, const ITab *__restrict__ itab_g,
              const int64_t total, const int work_bytes, const int use_tma, const int acc_path,
              const __grid_constant__ CUtensorMap smap, const __grid_constant__ CUtensorMap smap_nyq) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using K = PkInv<LOGMC, TF, NSUB>;
    const int tid = threadIdx.x;
    PkInvCtx c;
    c.a = args;
    c.work = reinterpret_cast<float *>(smem);
    c.ola = reinterpret_cast<float *>(smem);
    float *carry0 = reinterpret_cast<float *>(smem + work_bytes);
    // acc_path == 2: accumulator layout [work | landing][tables][accumulator] without carry buffers
    const int carry_elems = acc_path == 2 ? 0 : (args.N + 3) & ~3;
    float *carry1 = carry0 + carry_elems;
    int acc_off = work_bytes + 2 * 4 * carry_elems;
    if (K::STAB) acc_off += 4 * K::TWN + (int)sizeof(ITab) * K::MC;
    acc_off = (acc_off + 15) & ~15;
    if (K::STAB) {
        float *stw = carry1 + carry_elems;
        ITab *sitab = reinterpret_cast<ITab *>(stw + K::TWN);
        copy_table_f(stw, args.tw, K::TWN, tid, K::NT);
        copy_table_f(reinterpret_cast<float *>(sitab), reinterpret_cast<const float *>(itab_g), 4 * K::MC, tid, K::NT);
        c.tw = stw;
        c.itab = sitab;
    } else {
        c.tw = args.tw;
        c.itab = itab_g;
    }
    // TMA path (one-sided layouts, one pass-1 group pair per thread): the S tile of [Mc + 1 rows][TF
    // slices] lands OVER the work buffer -- one box [R1 rows][TF slices] per pass-1 group, row
    // j R1 + a = bin j + L1 a, plus the Nyquist row -- at 7 TB/s with no load instructions in the
    // warps (profiles/r1c_microbench_store_lds.txt); columns outside S arrive as zeros.
    __shared__ __align__(8) unsigned long long land_bar_mem;
    const unsigned land_bar = tma::smem_u32(&land_bar_mem);
    const bool tma_path = ONES && use_tma && K::WPT2 == 1;
    unsigned land_phase = 0;
    if (tma_path && tid == 0) {
        tma::mbar_init(land_bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    const int64_t lc = (int64_t)TF * args.chunk_tiles - args.halo;      // owned slices per chunk
    if (ONES && tma_path && acc_path) {
        // mfft 512, 75 % overlap: TMA-landed S + accumulator overlap-add.  The next tile's boxes are
        // requested as soon as the last pass has read the work buffer and land during the four
        // accumulator passes and the output of the current tile.
        float *acc = reinterpret_cast<float *>(smem + acc_off);
        auto land = [&](int64_t b, int64_t qa) { pk_land<K>(smap, smap_nyq, c.work, land_bar, tid, b, qa - args.p_min); };
        int64_t ch = blockIdx.x;
        int t = 0;
        bool have = ch < total;
        int64_t b = 0, qa = 0;
        if (have) {
            b = ch / args.cps;
            qa = args.qh0 + (ch - b * args.cps) * lc - args.halo;
            land(b, qa);
        }
        while (have) {
            // the tile after this one (same chunk, or the first tile of this CTA's next chunk)
            int64_t chn = ch, bn = b, qan = qa + TF;
            int tn = t + 1;
            bool have_n = true;
            if (tn >= args.chunk_tiles || qan > args.qh1) {
                chn = ch + gridDim.x;
                tn = 0;
                have_n = chn < total;
                if (have_n) {
                    bn = chn / args.cps;
                    qan = args.qh0 + (chn - bn * args.cps) * lc - args.halo;
                }
            }
            c.b = b;
            c.qa = qa;
            c.first = t == 0;
            c.land_shift = (int)((qa - args.p_min) & 1);
            typename K::Cp x0[K::R1], x1[K::R1], xn;
            tma::mbar_wait(land_bar, land_phase);
            land_phase ^= 1;
            K::p1p_load_smem(c, tid, 0, x0, x1, xn);
            __syncthreads();                   // every thread holds its bins before anyone overwrites the tile
            K::p1p_rest(c, tid, 0, x0, x1, xn);
            __syncthreads();
            if (K::NP == 3) {
                K::pass_mid(c, tid);
                __syncthreads();
            }
            {
                C2 r[K::GPT][K::RL];
                K::last_load(c, tid, r);
                __syncthreads();               // the work buffer is free from here on
                if (have_n) land(bn, qan);
                if (c.first) K::ola_zero_head(tid, acc);
                K::template ola_acc<3>(c, tid, r, acc);
                __syncthreads();
                K::template ola_acc<2>(c, tid, r, acc);
                __syncthreads();
                K::template ola_acc<1>(c, tid, r, acc);
                __syncthreads();
                K::template ola_acc<0>(c, tid, r, acc);
            }
            __syncthreads();
            K::ola_out(c, tid, acc);
            const typename K::Tail tail = K::ola_take_tail(tid, acc);
            __syncthreads();                   // head and tail read before the head is replaced
            K::ola_keep_tail(tid, acc, tail);
            ch = chn; t = tn; b = bn; qa = qan; have = have_n;
        }
        return;
    }
    // A CTA walks the tiles of one chunk of one signal in order (streaming overlap-add): only the
    // first tile of a chunk re-computes the halo slices, every other tile owns all TF slices.
    for (int64_t ch = blockIdx.x; ch < total; ch += gridDim.x) {
        c.b = ch / args.cps;
        const int64_t a0 = args.qh0 + (ch - c.b * args.cps) * lc;      // first owned slice
        int flip = 0;
#pragma unroll 1
        for (int t = 0; t < args.chunk_tiles; ++t) {
            c.qa = a0 - args.halo + (int64_t)t * TF;
            if (c.qa > args.qh1) break;
            c.first = t == 0;
            c.carry_in = flip ? carry1 : carry0;
            c.carry_out = flip ? carry0 : carry1;
            flip ^= 1;
            if (tma_path) {
                pk_land<K>(smap, smap_nyq, c.work, land_bar, tid, c.b, c.qa - args.p_min);
                c.land_shift = (int)((c.qa - args.p_min) & 1);
                typename K::Cp x0[K::R1], x1[K::R1], xn;
                tma::mbar_wait(land_bar, land_phase);
                land_phase ^= 1;
                K::p1p_load_smem(c, tid, 0, x0, x1, xn);
                __syncthreads();               // every thread holds its bins before anyone overwrites the tile
                K::p1p_rest(c, tid, 0, x0, x1, xn);
            } else if (ONES && !(args.dbg & 64)) {
#pragma unroll 1
                for (int i = 0; i < K::WPT2; ++i) {
#ifdef VEILS_ABLATE
                    if (args.dbg & 16) continue;
#endif
                    K::p1p(c, tid, i);
                }
            } else
#pragma unroll 1
            for (int i = 0; i < K::WPT; ++i) {
                Raw x[K::R1];
                Raw xn;
#ifdef VEILS_ABLATE
                if (args.dbg & 1) {            // ablation: no global loads
#pragma unroll
                    for (int a = 0; a < K::R1; ++a) { x[a].r0 = tid + a; x[a].i0 = a; x[a].r1 = tid; x[a].i1 = 1.f; }
                    xn = x[0];
                } else
#endif
                K::template p1_load_v<ONES>(c, tid, i, x, xn);
#ifdef VEILS_ABLATE
                if (args.dbg & 16) {           // ablation: no pass 1 arithmetic
                    if (x[3].r0 == 12345.678f) c.work[tid] = x[5].i1 + xn.r0;
                    continue;
                }
#endif
                K::p1_rest(c, tid, i, x, xn, nullptr);
            }
            __syncthreads();
            if (K::NP == 3) {
                K::pass_mid(c, tid);
                __syncthreads();
            }
#ifdef VEILS_ABLATE
            if (!(args.dbg & 8))
#endif
            {
                C2 r[K::GPT][K::RL];
                K::last_load(c, tid, r);
                __syncthreads();               // all reads of `work` done before it becomes `ola`
                K::last_store(c, tid, r);
            }
            __syncthreads();
#ifdef VEILS_ABLATE
            if (!(args.dbg & 4))
#endif
            K::gather(c, tid);
            __syncthreads();                   // `ola` consumed before the next tile overwrites `work`
        }
    }
}

#define pk_env_int(name, dflt) VEILS_ENV_ONCE(name, dflt)      // environment switches: read once per process
int pk_cta_cap(int per_sm) {
    const int e = VEILS_ENV_ONCE("VEILS_PK_CTAS_PER_SM", 0);
    return (e > 0 && e < per_sm) ? e : per_sm;
}

struct PkFwdLaunch {
    const float *x; void *out; const FwdParams *prm; const PTab *ptab; const float *tw; cudaStream_t st;
    template <int LOGMC, int TF, int NSUB>
    cudaError_t run() {
        const FwdParams &p = *prm;
        FwdArgs<float> a = make_fwd_args<float>(x, out, p, nullptr, tw, TF);
        a.dbg = pk_env_int("VEILS_PK_DBG", 0);
        const int xin_elems = (int)((fwd_xin_elems(p.N, p.H, TF, a.lay) + 3) & ~3LL);
        const size_t smem = pk_fwd_smem_bytes<LOGMC, TF>(p.N, xin_elems);
        using KernT = void (*)(const FwdArgs<float>, const PTab *, const int64_t, const int, const int, const int, const int,
                               const CUtensorMap, const CUtensorMap);
        const bool ones = p.mode >= 2, spec = p.spectrogram != 0;
        using K = PkFwd<LOGMC, TF, NSUB>;
        // one-sided layouts leave through the TMA engine when S can be described by a tensor map
        // (16-byte aligned base and row pitch); VEILS_PK_NO_TMA=1 forces the plain-store kernels
        CUtensorMap smap, smap_tile;
        memset(&smap, 0, sizeof(smap));
        memset(&smap_tile, 0, sizeof(smap_tile));
        bool stage = ones && p.batch < (1LL << 31) && !pk_env_int("VEILS_PK_NO_TMA", 0);
        if (stage) {
            const int esz = spec ? 4 : 8;
            const MapKey k1{out, {esz, p.P, p.ldp, p.F, p.batch, K::G * 1000 + K::RL, TF, 101}};
            stage = cached_map(&smap, k1, [&](CUtensorMap *m) {
                return tma::make_s_map(m, out, esz, p.P, p.ldp, p.F, p.batch, K::G, K::RL, TF, K::RL);
            });
            if (stage && !spec) {
                const MapKey k2{out, {8, p.P, p.ldp, p.F, p.batch, K::G * 1000 + K::RL, TF, 102 + K::NP}};
                stage = cached_map(&smap_tile, k2, [&](CUtensorMap *m) {
                    return K::NP == 2 ? tma::make_s_map(m, out, 8, p.P, p.ldp, p.F, p.batch, K::G, K::RL, TF, K::RL, K::G)
                                      : tma::make_s_map5(m, out, 8, p.P, p.ldp, p.F, p.batch, K::G, K::RL, K::R1, TF);
                });
            }
        }
        KernT kern = stage ? (spec ? (KernT)pk_fwd_kernel<LOGMC, TF, NSUB, true, true, true> : (KernT)pk_fwd_kernel<LOGMC, TF, NSUB, true, false, true>)
                   : ones  ? (spec ? (KernT)pk_fwd_kernel<LOGMC, TF, NSUB, true, true, false> : (KernT)pk_fwd_kernel<LOGMC, TF, NSUB, true, false, false>)
                           : (spec ? (KernT)pk_fwd_kernel<LOGMC, TF, NSUB, false, true, false> : (KernT)pk_fwd_kernel<LOGMC, TF, NSUB, false, false, false>);
        static KernelSlot slots[6];
        int per_sm = 0, dev = 0;
        cudaError_t e = prepare_kernel(slots[(stage ? 0 : (ones ? 2 : 4)) + (spec ? 1 : 0)], kern, (TF / 2) * NSUB, smem, &per_sm, &dev);
        if (e != cudaSuccess) return e;
        const int sms = device_sms(dev);
        const int64_t total = a.ptiles * p.batch;
        const int64_t cap = (int64_t)pk_cta_cap(per_sm) * sms;
        kern<<<(unsigned)(total < cap ? total : cap), (TF / 2) * NSUB, smem, st>>>(
            a, ptab, total, xin_elems, pk_zero_off(p.N, p.H, TF), pk_env_int("VEILS_PK_STAGGER_NS", 0), sms, smap, smap_tile);
        ++g_launch_count;
        return cudaGetLastError();
    }
};

struct PkInvLaunch {
    const void *S; float *xo; const InvParams *prm; const ITab *itab; const float *tw; cudaStream_t st;
    template <int LOGMC, int TF, int NSUB>
    cudaError_t run() {
        const InvParams &p = *prm;
        if (TF - inv_halo(p.N, p.H) < 1) return cudaErrorInvalidConfiguration;
        InvArgs<float> a = make_inv_args<float>(S, xo, p, nullptr, tw, TF);
        a.ostride = pk_ostride(p.N, p.ps);
        a.pair = pk_pair(p.N, p.ps) ? 1 : 0;
        a.dbg = pk_env_int("VEILS_PK_DBG_INV", 0);
        const int work_bytes = (int)pk_inv_work_bytes<LOGMC, TF>(p.N);
        const size_t smem = pk_inv_smem_bytes<LOGMC, TF>(p.N);
        using KernT = void (*)(const InvArgs<float>, const ITab *, const int64_t, const int, const int, const int,
                               const CUtensorMap, const CUtensorMap);
        KernT kern = p.mode >= 2 ? (KernT)pk_inv_kernel<LOGMC, TF, NSUB, true> : (KernT)pk_inv_kernel<LOGMC, TF, NSUB, false>;
        // one-sided layouts: S tiles land through the TMA engine when S can be described by a tensor map
        // (16-byte aligned base / row pitch) and the landed tile fits the work buffer
        using K = PkInv<LOGMC, TF, NSUB>;
        CUtensorMap smap, smap_nyq;
        memset(&smap, 0, sizeof(smap));
        memset(&smap_nyq, 0, sizeof(smap_nyq));
        // Measured (profiles/r1c_ablation.txt): landing alone (waited for at the start of the tile) pays
        // only for mfft = 512 (1.05 -> 0.96 ms; slower for mfft 256 and >= 2048); together with the
        // accumulator overlap-add the next tile lands during the current one's overlap-add.
        const bool acc_geo = a.halo == 3 && p.N == p.M && 4 * p.H == p.N && p.ps == p.N / 2 && a.pair &&
                             p.out_pitch % 2 == 0 && p.k0 % 2 == 0 && (((uintptr_t)xo) & 7) == 0 &&
                             !pk_env_int("VEILS_PK_NO_ACC", 0);
        // measured (one landing box per tile): mfft 256 1.21 -> 0.98 ms, mfft 512 0.96 -> 0.77 ms,
        // mfft 2048 1.32 -> 1.14 ms, mfft 4096 2.66 -> 2.55 ms (mfft 1024 has two group pairs per thread)
        const bool acc_size = (LOGMC == 7 || LOGMC == 8 || LOGMC == 10 || LOGMC == 11) &&
                              pk_inv_acc_smem_bytes<LOGMC, TF>(p.N) + 1024 <= SMEM_LIMIT;
        bool use_tma = (LOGMC == 8 || (acc_geo && acc_size)) && (K::R1 * K::LTF * 8) % 128 == 0 &&
                       p.mode >= 2 && K::WPT2 == 1 && p.batch < (1LL << 31) && p.P < (1LL << 31) &&
                       !pk_env_int("VEILS_PK_NO_TMA", 0) && !pk_env_int("VEILS_PK_NO_TMA_LOAD", 0) && K::L1 <= 256;
        if (use_tma) {
            const MapKey k1{S, {8, p.P, p.ldp, p.F, p.batch, K::L1 * 1000 + K::R1, K::LTF, 201}};
            use_tma = cached_map(&smap, k1, [&](CUtensorMap *m) {
                return tma::make_s_map(m, S, 8, p.P, p.ldp, p.F, p.batch, K::L1, K::R1, K::LTF, K::R1, K::L1);
            });
            if (use_tma) {
                const MapKey k2{S, {8, p.P, p.ldp, p.F, p.batch, K::L1 * 1000 + K::R1, K::LTF, 202}};
                use_tma = cached_map(&smap_nyq, k2, [&](CUtensorMap *m) {
                    return tma::make_s_map(m, S, 8, p.P, p.ldp, p.F, p.batch, K::L1, K::R1, K::LTF, 1);
                });
            }
        }
        // accumulator overlap-add: 75 % overlap with the default roll, 8-byte aligned output pairs.
        // 2 = its own shared-memory layout (no carry buffers); the plain TMA path keeps the base layout
        const int acc_path = (use_tma && acc_geo && acc_size) ? 2 : 0;
        int work_bytes_l = work_bytes;
        size_t smem_l = smem;
        if (acc_path) {
            work_bytes_l = (int)pk_inv_acc_work_bytes<LOGMC, TF>(p.N);
            smem_l = pk_inv_acc_smem_bytes<LOGMC, TF>(p.N);
        }
        const int use_tma_l = use_tma && (acc_path || (size_t)work_bytes >= K::LAND_BYTES);
        static KernelSlot slots[2];
        int per_sm = 0, dev = 0;
        cudaError_t e = prepare_kernel(slots[p.mode >= 2 ? 1 : 0], kern, (TF / 2) * NSUB, smem_l, &per_sm, &dev);
        if (e != cudaSuccess) return e;
        const int64_t cap = (int64_t)pk_cta_cap(per_sm) * device_sms(dev);
        pk_inv_chunks(a, p, TF, p.batch, cap);
        const int64_t total = a.cps * p.batch;
        kern<<<(unsigned)(total < cap ? total : cap), (TF / 2) * NSUB, smem_l, st>>>(a, itab, total, work_bytes_l, use_tma_l, acc_path, smap, smap_nyq);
        ++g_launch_count;
        return cudaGetLastError();
    }
};

}  // namespace

bool pk_supported(int64_t M, int64_t N, int64_t H, int64_t ps, bool inverse) {
    return p2::pk_supported(M, N, H, ps, inverse);
}
int64_t pk_twiddle_count(int64_t M) { return p2::pk_twiddle_count(M); }
void pk_itab_fill_host(int64_t M, int64_t N, int64_t ps, const double *dual_scaled, void *out16) {
    p2::pk_itab_fill(M, N, ps, dual_scaled, (p2::ITab *)out16);
}
void pk_ptab_fill_host(int64_t M, int64_t N, int64_t H, int64_t ps, const double *win, void *out16) {
    p2::Cfg c;
    if (!p2::pk_cfg_for(M, &c)) return;
    p2::pk_ptab_fill(M, N, H, ps, win, c.tf, (p2::PTab *)out16);
}
void pk_twiddle_fill(int64_t M, bool inverse, double *t) { p2::pk_twiddle_fill(M, inverse, t); }

cudaError_t pk_forward(const float *x, void *out, const FwdParams &prm, const Tables<float> &tb, cudaStream_t st) {
    if (prm.P <= 0 || prm.batch <= 0) return cudaSuccess;
    p2::Cfg c;
    if (!p2::pk_cfg_for(prm.M, &c)) return cudaErrorNotSupported;
    PkFwdLaunch l{x, out, &prm, (const PTab *)tb.ptab_pk, tb.tw_pk, st};
    return p2::pk_dispatch(c.logmc, l);
}

cudaError_t pk_inverse(const void *S, float *x_out, const InvParams &prm, const Tables<float> &tb, cudaStream_t st) {
    if (prm.batch <= 0 || prm.k1 <= prm.k0) return cudaSuccess;
    p2::Cfg c;
    if (!p2::pk_cfg_for(prm.M, &c)) return cudaErrorNotSupported;
    PkInvLaunch l{S, x_out, &prm, (const ITab *)tb.itab_pk, tb.tw_pk_inv, st};
    return p2::pk_dispatch(c.logmc, l);
}

}  // namespace veils
