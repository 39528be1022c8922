#include "weights.h"

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "common.h"

namespace msb {

namespace {
template <typename T>
T read_le(const uint8_t* bytes, size_t size, size_t& pos) {
  if (pos + sizeof(T) > size) throw std::runtime_error("MSW: truncated header");
  T v;
  std::memcpy(&v, bytes + pos, sizeof(T));
  pos += sizeof(T);
  return v;
}
}  // namespace

void parse_msw(const uint8_t* bytes, size_t size, WeightFile& out) {
  if (bytes == nullptr || size < 16 || std::memcmp(bytes, "MSW1", 4) != 0) {
    throw std::runtime_error("MSW: bad magic (expected an MSW1 weight container)");
  }
  size_t pos = 4;
  uint32_t version = read_le<uint32_t>(bytes, size, pos);
  if (version != 1) throw std::runtime_error(format("MSW: unsupported version %u", version));
  out.arch = read_le<uint32_t>(bytes, size, pos);
  uint32_t n = read_le<uint32_t>(bytes, size, pos);
  for (uint32_t i = 0; i < n; i++) {
    uint16_t name_len = read_le<uint16_t>(bytes, size, pos);
    if (pos + name_len > size) throw std::runtime_error("MSW: truncated tensor name");
    std::string name(reinterpret_cast<const char*>(bytes + pos), name_len);
    pos += name_len;
    uint8_t dtype = read_le<uint8_t>(bytes, size, pos);
    uint8_t ndim = read_le<uint8_t>(bytes, size, pos);
    if (dtype != 0) throw std::runtime_error("MSW: only f32 tensors are supported");
    HostTensor t;
    size_t count = 1;
    for (int d = 0; d < ndim; d++) {
      uint32_t dim = read_le<uint32_t>(bytes, size, pos);
      t.shape.push_back(dim);
      if (dim != 0 && count > (SIZE_MAX / sizeof(float)) / dim) {
        throw std::runtime_error("MSW: tensor '" + name + "' has an element count that overflows");
      }
      count *= dim;
    }
    uint64_t off = read_le<uint64_t>(bytes, size, pos);
    uint64_t nbytes = read_le<uint64_t>(bytes, size, pos);
    if (nbytes != count * sizeof(float) || off > size || nbytes > size - off || (off % 4) != 0) {
      throw std::runtime_error("MSW: tensor '" + name + "' has an invalid extent");
    }
    t.data = reinterpret_cast<const float*>(bytes + off);
    t.count = count;
    out.tensors[name] = t;
  }
}

void load_msw_file(const std::string& path, WeightFile& out) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("Failed to open weight file '" + path + "'");
  std::streamsize sz = f.tellg();
  f.seekg(0);
  out.owned.resize((size_t)sz);
  if (!f.read(reinterpret_cast<char*>(out.owned.data()), sz)) {
    throw std::runtime_error("Failed to read weight file '" + path + "'");
  }
  parse_msw(out.owned.data(), out.owned.size(), out);
}

const HostTensor& WeightFile::get(const std::string& name) const {
  auto it = tensors.find(name);
  if (it == tensors.end()) throw std::runtime_error("Missing weight tensor '" + name + "'");
  return it->second;
}

const HostTensor& WeightFile::get(const std::string& name,
                                  std::initializer_list<int64_t> shape) const {
  const HostTensor& t = get(name);
  if (t.shape != std::vector<int64_t>(shape)) {
    throw std::runtime_error("Weight tensor '" + name + "' has an unexpected shape");
  }
  return t;
}

}  // namespace msb
