// pow2_tile.cu -- power-of-two tile family (placeholder until the fast path lands).
#include "kernels.cuh"
namespace veils {
bool pow2_supported(int64_t, int64_t, int64_t, int) { return false; }
int64_t pow2_twiddle_count(int64_t) { return 0; }
void pow2_twiddle_fill(int64_t, double *) {}
template <typename T>
cudaError_t pow2_forward(const T *, void *, const FwdParams &, const Tables<T> &, cudaStream_t) {
    return cudaErrorNotSupported;
}
template <typename T>
cudaError_t pow2_inverse(const void *, T *, const InvParams &, const Tables<T> &, cudaStream_t) {
    return cudaErrorNotSupported;
}
template cudaError_t pow2_forward<float>(const float *, void *, const FwdParams &, const Tables<float> &, cudaStream_t);
template cudaError_t pow2_forward<double>(const double *, void *, const FwdParams &, const Tables<double> &, cudaStream_t);
template cudaError_t pow2_inverse<float>(const void *, float *, const InvParams &, const Tables<float> &, cudaStream_t);
template cudaError_t pow2_inverse<double>(const void *, double *, const InvParams &, const Tables<double> &, cudaStream_t);
}  // namespace veils
