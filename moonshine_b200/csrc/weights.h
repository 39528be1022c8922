// Host-side weight container (.msw) reader.  Tensor names are the HF
// MoonshineForConditionalGeneration state-dict keys.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace msb {

struct HostTensor {
  std::vector<int64_t> shape;
  const float* data = nullptr;  // points into the container bytes
  size_t count = 0;
};

struct WeightFile {
  uint32_t arch = 0;
  std::map<std::string, HostTensor> tensors;
  // Keeps file bytes alive when loaded from disk (empty when the caller owns
  // the memory, as with moonshine_load_transcriber_from_memory_files).
  std::vector<uint8_t> owned;

  const HostTensor& get(const std::string& name) const;
  const HostTensor& get(const std::string& name, std::initializer_list<int64_t> shape) const;
};

// Parses an MSW1 container; throws std::runtime_error on malformed input.
void parse_msw(const uint8_t* bytes, size_t size, WeightFile& out);
void load_msw_file(const std::string& path, WeightFile& out);

}  // namespace msb
