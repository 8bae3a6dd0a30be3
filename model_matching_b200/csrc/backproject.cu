// backproject.cu -- depth image -> organised point cloud (reference src/rgbd.cpp:208-225).
// One thread per pixel; fp32 sub/mul/div in the reference's order, so results are bit-exact.
// 21 algorithmic bytes per pixel (2 depth + 3 BGR in, 12 xyz + 4 rgb out): launch/PCIe bound.
#include "stocs_ctx.h"

namespace {
__global__ void backproject_kernel(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ bgr,
                                   int W, int H, float fx, float cx, float fy, float cy, float scale,
                                   float* __restrict__ xyz, uint32_t* __restrict__ rgb) {
  const int n = W * H;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int i = k / W, j = k - i * W;
    const float d = (float)depth[k] * scale;
    xyz[3 * (size_t)k + 0] = ((float)j - cx) * d / fx;
    xyz[3 * (size_t)k + 1] = ((float)i - cy) * d / fy;
    xyz[3 * (size_t)k + 2] = d;
    if (rgb && bgr)
      rgb[k] = ((uint32_t)bgr[3 * (size_t)k + 2] << 16) | ((uint32_t)bgr[3 * (size_t)k + 1] << 8) |
               (uint32_t)bgr[3 * (size_t)k + 0];
  }
}
}  // namespace

int stocs_launch_backproject(stocs_b200_ctx* ctx, const uint16_t* d_depth, const uint8_t* d_bgr, int W,
                             int H, float fx, float cx, float fy, float cy, float scale, float* d_xyz,
                             uint32_t* d_rgb, cudaStream_t st) {
  const int n = W * H;
  int blocks = (n + 255) / 256;
  int cap = ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  backproject_kernel<<<blocks, 256, 0, st>>>(d_depth, d_bgr, W, H, fx, cx, fy, cy, scale, d_xyz, d_rgb);
  STOCS_CUDA(ctx, cudaGetLastError());
  return STOCS_OK;
}
