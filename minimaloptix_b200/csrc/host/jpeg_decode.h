// jpeg_decode.h — JPEG reader for albedo textures.  The reference loads textures
// through QImage (MinimalOptiX.cpp:445-479), i.e. through libjpeg; this decoder restates libjpeg's
// published arithmetic (integer "islow" inverse DCT, triangle-filter chroma upsampling, 16-bit
// fixed-point YCbCr -> RGB) so the texels equal what QImage hands to the reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace moxh {

// 8-bit baseline, extended-sequential and progressive Huffman JPEG with 1 (grey) or 3 components.
// Output: rows top-down, 3 bytes per pixel.
bool decodeJpeg(const uint8_t* data, size_t size, int& w, int& h, std::vector<uint8_t>& rgb, std::string& err);

}  // namespace moxh
