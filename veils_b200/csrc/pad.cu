// pad.cu -- border extension of the input signal for the twin's padding modes
// (python/standalone_stft.py:309-336: 'zeros' | 'edge' | 'even' | 'odd'; the reference's Rust
// crate only has zeros, src/lib.rs:631-645).  The extended signal
//     x_ext[i] = x[s]          0 <= s < n            (s = i - left)
//              = pad(s)        otherwise
//   edge: x[clamp(s)]      even: x[min(-s, n-1)] (s < 0), x[clamp(2n-2-s)] (s >= n)      odd: -even
// is then transformed by the ordinary zero-padding kernels with k_offset + left: every slice sees
// exactly the samples the twin's _x_slices would hand it.
#include "kernels.cuh"

namespace veils {

namespace {

template <typename T>
__global__ void pad_kernel(const T *__restrict__ x, T *__restrict__ xe, int64_t n, int64_t x_pitch, int64_t left,
                           int64_t n_ext, int64_t ext_pitch, int mode) {
    const int64_t b = blockIdx.y;
    const T *xb = x + b * x_pitch;
    T *eb = xe + b * ext_pitch;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ext; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i - left;
        T v;
        if (s >= 0 && s < n) {
            v = xb[s];
        } else if (mode == 0) {
            v = (T)0;
        } else if (mode == 1) {
            v = xb[s < 0 ? 0 : n - 1];
        } else {
            int64_t idx = s < 0 ? -s : 2 * n - 2 - s;
            idx = idx < 0 ? 0 : (idx > n - 1 ? n - 1 : idx);
            v = mode == 2 ? xb[idx] : -xb[idx];
        }
        eb[i] = v;
    }
}

}  // namespace

template <typename T>
cudaError_t pad_signal(const T *x, T *x_ext, int64_t n, int64_t batch, int64_t x_pitch, int64_t left, int64_t n_ext,
                       int64_t ext_pitch, int mode, cudaStream_t st) {
    if (batch <= 0 || n_ext <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t bx = (n_ext + threads * 4 - 1) / (threads * 4);
    if (bx > 65535) bx = 65535;
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        pad_kernel<T><<<dim3((unsigned)bx, (unsigned)nb), threads, 0, st>>>(x + b0 * x_pitch, x_ext + b0 * ext_pitch, n,
                                                                             x_pitch, left, n_ext, ext_pitch, mode);
        ++g_launch_count;
    }
    return cudaGetLastError();
}
template cudaError_t pad_signal<float>(const float *, float *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);
template cudaError_t pad_signal<double>(const double *, double *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);

}  // namespace veils
