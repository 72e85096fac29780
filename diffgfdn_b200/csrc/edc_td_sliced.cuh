// Interface between the receiver-step entry point (edc_td_fused.cu: dgfdn_td_edc_fused) and the time-sliced
// persistent kernel K3t (edc_td_sliced.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgfdn {

constexpr int kSlicedNsp = 160;                    // padded slice count of a row record (>= #SMs, multiple of 32)
constexpr int kSlicedRecFloats = 6 * kSlicedNsp;   // per-row record: T1 | T2 (floats) | part_gs (float4 per slice)
constexpr int kSlicedHeaderBytes = 4096;           // part_loss[kSlicedNsp] doubles

struct SlicedShape {
  int shape;              // index into the (lanes per row, float4 per lane) table
  int ns;                 // slices = CTAs of the launch
  int samples_per_slice;
  int threads;
};

// Shape the sliced kernel would use for (g, tn) on the current device; false when it does not take the shape.
bool sliced_shape(int g, int64_t tn, SlicedShape* out);
size_t sliced_ws_bytes(int64_t rows);
// The carry words of the workspace must read "not published" (0xFF..) before the first launch; every launch leaves
// them that way.
int sliced_ws_init(void* ws, int g, int64_t rows, int64_t tn, cudaStream_t st);
int sliced_launch(int g, int64_t rows, int64_t tn, const float* s, const float* hy, const float* hd, int64_t ldhd,
                  const float* target_db, int64_t ldt, const float* mask, double coef, double* loss_sum, float* gs, float* ghy,
                  int accumulate, void* ws, cudaStream_t st);

}  // namespace dgfdn
