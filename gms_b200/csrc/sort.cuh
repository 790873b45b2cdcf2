// sort.cuh — device-wide sort / scan primitives used by the graph builders.
// The LSD radix sort and scans come from CUB (ships with the CUDA toolkit); they sit outside the timed
// intersection kernels.  Every helper runs on the library stream and owns its temp storage.
#pragma once
#include <cub/cub.cuh>
#include "common.cuh"

namespace gmsb {

// Sorts `n` 64-bit keys on bits [begin_bit, end_bit); returns the buffer (keys or alt) holding the result.
inline uint64_t *radix_sort_keys(uint64_t *keys, uint64_t *alt, int64_t n, int begin_bit, int end_bit,
                                 bool descending = false) {
    cub::DoubleBuffer<uint64_t> db(keys, alt);
    size_t bytes = 0;
    cudaStream_t s = rt().stream;
    end_bit = end_bit > 64 ? 64 : end_bit;
    if (descending) GMSB_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, bytes, db, n, begin_bit, end_bit, s));
    else GMSB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, db, n, begin_bit, end_bit, s));
    DevBuf<uint8_t> tmp(bytes);
    if (descending) GMSB_CUDA(cub::DeviceRadixSort::SortKeysDescending(tmp.p, bytes, db, n, begin_bit, end_bit, s));
    else GMSB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, db, n, begin_bit, end_bit, s));
    rt().launches += (uint64_t)((end_bit - begin_bit + 7) / 8 + 1);   // onesweep: histogram + one pass per digit
    GMSB_CUDA(cudaStreamSynchronize(s));
    return db.Current();
}

// Stable sort of (key32, value64) pairs on key bits [0, end_bit); results land in *keys_out / *vals_out.
inline void radix_sort_pairs(uint32_t *keys, uint32_t *keys_alt, uint64_t *vals, uint64_t *vals_alt, int64_t n,
                             int end_bit, uint32_t **keys_out, uint64_t **vals_out) {
    cub::DoubleBuffer<uint32_t> dk(keys, keys_alt);
    cub::DoubleBuffer<uint64_t> dv(vals, vals_alt);
    size_t bytes = 0;
    cudaStream_t s = rt().stream;
    end_bit = end_bit > 32 ? 32 : end_bit;
    GMSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, 0, end_bit, s));
    DevBuf<uint8_t> tmp(bytes);
    GMSB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, dk, dv, n, 0, end_bit, s));
    rt().launches += (uint64_t)((end_bit + 7) / 8 + 1);
    GMSB_CUDA(cudaStreamSynchronize(s));
    *keys_out = dk.Current();
    *vals_out = dv.Current();
}

// Stable sort of (key32, value32) pairs on key bits [0, end_bit); returns the buffer holding the sorted values
// (vals or vals_alt).
inline int32_t *radix_sort_pairs32(uint32_t *keys, uint32_t *keys_alt, int32_t *vals, int32_t *vals_alt, int64_t n,
                                   int end_bit) {
    cub::DoubleBuffer<uint32_t> dk(keys, keys_alt);
    cub::DoubleBuffer<int32_t> dv(vals, vals_alt);
    size_t bytes = 0;
    cudaStream_t s = rt().stream;
    end_bit = end_bit > 32 ? 32 : end_bit;
    GMSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, 0, end_bit, s));
    DevBuf<uint8_t> tmp(bytes);
    GMSB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, dk, dv, n, 0, end_bit, s));
    rt().launches += (uint64_t)((end_bit + 7) / 8 + 1);
    GMSB_CUDA(cudaStreamSynchronize(s));
    return dv.Current();
}

template <typename In, typename Out>
inline void exclusive_sum(const In *in, Out *out, int64_t n) {
    size_t bytes = 0;
    cudaStream_t s = rt().stream;
    auto it = cub::TransformInputIterator<Out, cub::CastOp<Out>, const In *>(in, cub::CastOp<Out>());
    GMSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, n, s));
    DevBuf<uint8_t> tmp(bytes);
    GMSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, it, out, n, s));
    rt().launches += 2;
    GMSB_CUDA(cudaStreamSynchronize(s));
}

template <typename T>
inline void inclusive_sum_inplace(T *data, int64_t n) {
    size_t bytes = 0;
    cudaStream_t s = rt().stream;
    GMSB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, data, data, n, s));
    DevBuf<uint8_t> tmp(bytes);
    GMSB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, data, data, n, s));
    rt().launches += 2;
    GMSB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace gmsb
