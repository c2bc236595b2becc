#pragma once
#include <cstdint>
#include <string>
namespace moxh {
void accumToRgb8(const float* accum, uint32_t W, uint32_t H, float nAccum, uint8_t* out);
bool writeImage(const std::string& path, const uint8_t* rgb, uint32_t W, uint32_t H, std::string& err);
bool writeAccum(const std::string& path, const float* accum, uint32_t W, uint32_t H, uint64_t launches, std::string& err);
bool readAccum(const std::string& path, float* accum, uint32_t W, uint32_t H, uint64_t* launches, std::string& err);
}
