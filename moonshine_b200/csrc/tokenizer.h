// tokenizer.bin reader + detokeniser.  Semantics follow the reference's
// BinTokenizer (core/bin-tokenizer/bin-tokenizer.cpp:46-66 record format,
// :406-425 tokens_to_text); written from that behaviour contract.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace msb {

class Tokenizer {
 public:
  Tokenizer(const uint8_t* data, size_t size);
  static Tokenizer* from_file(const std::string& path);
  size_t size() const { return pieces_.size(); }
  // Throws std::out_of_range / std::runtime_error like the reference's
  // `.at(token)` and empty-token check.
  std::string tokens_to_text(const std::vector<int32_t>& tokens, bool skip_specials = true) const;
  // piece begins with the word-boundary marker U+2581 (core/word-alignment.cpp:158-170); false for bad ids
  bool starts_word(int32_t token) const;

 private:
  std::vector<std::string> pieces_;
};

// Replace structurally invalid UTF-8 by '?' (core/transcriber.cpp:1489-1541).
std::string sanitize_utf8(const std::string& text);

}  // namespace msb
