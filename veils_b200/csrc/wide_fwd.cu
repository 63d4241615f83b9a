// wide_fwd.cu -- forward launcher of the wide family (mfft 1024 / 2048 / 4096, default geometry).
#include "wide_launch.cuh"

namespace veils {

namespace {
using namespace wd;

struct FwdLaunch {
    const float *x; void *out; const FwdParams *prm; const float *tab; cudaStream_t st;
    template <int LOGMC, int TF, int MODE, bool SPEC>
    cudaError_t go(const WFwdArgs &a) {
        using K = WFwd<LOGMC, TF>;
        auto kern = wide_fwd_kernel<LOGMC, TF, MODE, SPEC>;
        const size_t smem = wide_fwd_smem_bytes<LOGMC, TF>();
        int cap = 0;
        static KernelSlot slot;
        cudaError_t e = wide_prepare(slot, kern, K::NT, smem, &cap);
        if (e != cudaSuccess) return e;
        const int64_t total = a.ptiles * prm->batch;
        kern<<<(unsigned)(total < cap ? total : cap), K::NT, smem, st>>>(a, total);
        ++g_launch_count;
        return cudaGetLastError();
    }
    template <int LOGMC, int TF>
    cudaError_t run() {
        using K = WFwd<LOGMC, TF>;
        const FwdParams &p = *prm;
        const WTabOff o = wide_table_offsets(p.M);
        const fl2 *t = reinterpret_cast<const fl2 *>(tab);
        WFwdArgs a;
        a.x = x; a.out = out;
        a.tab = t;
        (void)o;
        a.n = p.n; a.x_pitch = p.x_pitch; a.ldp = p.ldp; a.p0 = p.p0; a.P = p.P; a.k_offset = p.k_offset;
        a.ptiles = (p.P + TF - 1) / TF; a.F = p.F;
        a.mode = p.mode; a.spectrogram = p.spectrogram;
        const int esz = p.spectrogram ? 4 : 8;
        a.vec16 = ((p.ldp * esz) % 16 == 0 && ((uintptr_t)out % 16) == 0) ? 1 : 0;
        a.fac2x = p.mode == 3 ? (float)p.fac2x : 1.f;
        a.no_bulk = env_int("VEILS_WIDE_NO_BULK", 0);
        (void)sizeof(K);
        if (p.spectrogram) {
            switch (p.mode) {
                case 0: return go<LOGMC, TF, 0, true>(a);
                case 1: return go<LOGMC, TF, 1, true>(a);
                case 2: return go<LOGMC, TF, 2, true>(a);
                default: return go<LOGMC, TF, 3, true>(a);
            }
        }
        switch (p.mode) {
            case 0: return go<LOGMC, TF, 0, false>(a);
            case 1: return go<LOGMC, TF, 1, false>(a);
            case 2: return go<LOGMC, TF, 2, false>(a);
            default: return go<LOGMC, TF, 3, false>(a);
        }
    }
};
}  // namespace

bool wide_supported(int64_t M, int64_t N, int64_t H, int64_t ps) { return wd::wide_geometry_ok(M, N, H, ps); }
int64_t wide_table_floats(int64_t M) { return 2 * wd::wide_table_entries(M); }
// slices per tile of the plan (VEILS_WIDE_TF is a tuning switch, read at PLAN creation)
int wide_tf(int64_t M) {
    wd::WCfg c;
    const char *e = getenv("VEILS_WIDE_TF");
    if (!wd::wide_cfg_for(M, e ? atoi(e) : 0, &c)) return 0;
    return c.tf;
}
// forward: window * 0.5; inverse: dual * scale.  `out` receives wide_table_floats(M) doubles.
void wide_table_fill_host(int64_t M, int tf, bool inverse, const double *vec_n, double scale, double *out) {
    wd::WCfg c;
    if (!wd::wide_cfg_for(M, tf, &c)) return;
    int r1 = 0, r2 = 0;
    if (c.logmc == 9) { r1 = 8; r2 = 8; }
    else if (c.logmc == 10 && c.tf == 16) { r1 = inverse ? 8 : 16; r2 = 8; }
    else if (c.logmc == 10) { r1 = 8; r2 = 16; }
    else { r1 = inverse ? 8 : 16; r2 = 16; }
    wd::wide_table_fill(M, r1, r2, vec_n, scale, out);
}

cudaError_t wide_forward(const float *x, void *out, const FwdParams &prm, const float *tab, int tf, cudaStream_t st) {
    if (prm.P <= 0 || prm.batch <= 0) return cudaSuccess;
    wd::WCfg c;
    if (!wd::wide_cfg_for(prm.M, tf, &c)) return cudaErrorNotSupported;
    FwdLaunch l{x, out, &prm, tab, st};
    return wd::wide_dispatch(c, l);
}

}  // namespace veils
