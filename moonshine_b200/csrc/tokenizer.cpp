#include "tokenizer.h"

#include <fstream>
#include <stdexcept>

namespace msb {

Tokenizer::Tokenizer(const uint8_t* data, size_t size) {
  if (data == nullptr || size == 0) throw std::runtime_error("Tokenizer data is nullptr or empty");
  size_t p = 0;
  while (p < size) {
    uint8_t b0 = data[p++];
    if (b0 == 0) {           // empty record keeps ids aligned
      pieces_.emplace_back();
      continue;
    }
    size_t len;
    if (b0 < 128) {
      len = b0;
    } else {                 // two-byte length: b1 * 128 + b0 - 128
      if (p >= size) throw std::runtime_error("Truncated tokenizer data: missing length byte");
      len = (size_t)data[p++] * 128 + b0 - 128;
    }
    if (len > size - p) throw std::runtime_error("Truncated tokenizer data: token exceeds input");
    pieces_.emplace_back(reinterpret_cast<const char*>(data + p), len);
    p += len;
  }
  if (pieces_.empty()) throw std::runtime_error("No tokens found in tokenizer data");
}

Tokenizer* Tokenizer::from_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("Failed to open tokenizer file at " + path);
  std::streamsize sz = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> buf((size_t)sz);
  if (sz > 0 && !f.read(reinterpret_cast<char*>(buf.data()), sz)) {
    throw std::runtime_error("Failed to read tokenizer file at " + path);
  }
  return new Tokenizer(buf.data(), buf.size());
}

std::string Tokenizer::tokens_to_text(const std::vector<int32_t>& tokens, bool skip_specials) const {
  static const std::string kSpace = "\xE2\x96\x81";  // U+2581
  std::string bytes;
  for (int32_t t : tokens) {
    const std::string& piece = pieces_.at((size_t)t);
    if (piece.empty()) throw std::runtime_error("Invalid token " + std::to_string(t));
    if (skip_specials && piece.size() > 2 && piece.front() == '<' && piece.back() == '>') continue;
    bytes += piece;
  }
  std::string out;
  out.reserve(bytes.size());
  for (size_t i = 0; i < bytes.size();) {
    if (bytes.compare(i, kSpace.size(), kSpace) == 0) {
      out.push_back(' ');
      i += kSpace.size();
    } else {
      out.push_back(bytes[i++]);
    }
  }
  auto is_ws = [](unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
  size_t b = 0, e = out.size();
  while (b < e && is_ws((unsigned char)out[b])) b++;
  while (e > b && is_ws((unsigned char)out[e - 1])) e--;
  return out.substr(b, e - b);
}

bool Tokenizer::starts_word(int32_t token) const {
  if (token < 0 || (size_t)token >= pieces_.size()) return false;
  const std::string& p = pieces_[(size_t)token];
  return p.size() >= 3 && (uint8_t)p[0] == 0xE2 && (uint8_t)p[1] == 0x96 && (uint8_t)p[2] == 0x81;
}

std::string sanitize_utf8(const std::string& s) {
  std::string out;
  out.reserve(s.size());
  auto cont = [&](size_t k) { return ((uint8_t)s[k] & 0xC0) == 0x80; };
  size_t i = 0, n = s.size();
  while (i < n) {
    uint8_t c = (uint8_t)s[i];
    size_t need = c < 0x80 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : (c & 0xF8) == 0xF0 ? 4 : 0;
    bool ok = need != 0 && n - i >= need;
    for (size_t k = 1; ok && k < need; k++) ok = cont(i + k);
    if (!ok) {
      out.push_back('?');
      i++;
    } else {
      out.append(s, i, need);
      i += need;
    }
  }
  return out;
}

}  // namespace msb
