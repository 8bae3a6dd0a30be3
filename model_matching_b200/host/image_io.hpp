// image_io.hpp -- minimal PNG reader/writer (zlib) standing in for cv::imread / cv::imwrite
// (reference src/rgbd.cpp:197-199, src/stocs.cpp:116,625): 8-bit gray / RGB / RGBA and 16-bit
// gray, non-interlaced.  OpenCV is not available in this image.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace imgio {

struct Image {
  int width = 0, height = 0, channels = 0, bit_depth = 0;  // channels as stored (1, 3 or 4)
  std::vector<uint8_t> u8;    // bit_depth 8: height*width*channels, RGB(A) order
  std::vector<uint16_t> u16;  // bit_depth 16: height*width (gray only)
  bool empty() const { return width == 0; }
};

bool read_png(const std::string& path, Image& out, std::string* err = nullptr);
bool write_png_gray8(const std::string& path, const uint8_t* data, int width, int height);

// cv::imread(path, CV_LOAD_IMAGE_COLOR): 8-bit BGR, H*W*3
bool load_bgr8(const std::string& path, std::vector<uint8_t>& bgr, int& W, int& H);
// cv::imread(path, CV_16UC1) as the reference uses it for depth / probability maps: 16-bit gray
bool load_gray16(const std::string& path, std::vector<uint16_t>& g, int& W, int& H);
// cv::imread(path, CV_8UC1) for the edge map: 8-bit gray
bool load_gray8(const std::string& path, std::vector<uint8_t>& g, int& W, int& H);

}  // namespace imgio
