// tokenizer.bin reader + detokeniser.  Semantics follow the reference's
// BinTokenizer (core/bin-tokenizer/bin-tokenizer.cpp:46-66 record format,
// :406-425 tokens_to_text); written from that behaviour contract.
#pragma once
#include <unordered_map>
#include <cstdint>
#include <string>
#include <utility>
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
  // Text -> ids (reference: BinTokenizer::text_to_tokens, core/bin-tokenizer/bin-tokenizer.cpp:274-404).
  // bpe = true: byte-pair encoding with the id as merge rank (what the streaming models use,
  // core/moonshine-streaming-model.cpp:57); falls back to longest match when the vocabulary has no
  // 256-entry byte block.  bpe = false: greedy longest match, ties to the lowest id (throws on no match).
  std::vector<int32_t> text_to_tokens(const std::string& text, bool bpe) const;

 private:
  void build_encoder_index();
  std::vector<int32_t> encode_longest_match(const std::string& text) const;
  std::vector<std::string> pieces_;
  std::vector<std::vector<int32_t>> by_first_byte_;        // ascending ids per first byte
  std::unordered_map<std::string, int32_t> merge_ids_;     // piece -> lowest id, pieces after the byte block
  int32_t byte_base_ = -1;                                 // first id of the 0x00..0xFF block, -1 if absent
};

// Key terms out of a free-form passage (reference: ContextExtractor::extract, core/context-extractor.cpp:186-251):
// words of >= 3 characters without digits, possessives stripped, case variants grouped (most frequent spelling
// wins), kept when the tokenizer needs >= 2 subwords for " word", ranked by occurrences, then subwords, then
// first appearance; at most max_terms (<= 0: 200).
std::vector<std::string> extract_key_terms(const std::string& context, int32_t max_terms, const Tokenizer& tokenizer);

// Key-term biasing (reference: ContextBiaser, core/context-biaser.{h,cpp}): a trie over the token spellings
// of the key terms; before each argmax the logits of the tokens that continue an active path get
// boost * (1 + ln(depth)), the largest bonus when several paths propose the same token.
class KeytermBiaser {
 public:
  void clear();
  void set_boost(float boost) { boost_ = boost; }
  bool empty() const { return sequences_ == 0; }
  void add_token_sequence(const std::vector<int32_t>& tokens);
  static std::vector<std::string> variants_for_term(const std::string& term);
  // Walk state is per decoded sequence: a batch keeps one Walk per utterance over the shared trie.
  struct Walk {
    std::vector<int32_t> active{0};
  };
  void apply(const Walk& w, float* logits, int vocab) const;
  void advance(Walk& w, int32_t token) const;
  // The same bonuses in sparse form (on-device biasing): the root is always active, so its children's bonus holds
  // for every utterance at every step (dense [vocab]); the deeper active nodes of a walk add, per token, the
  // difference between the merged (largest) bonus apply() would use and the root's share.
  std::vector<float> root_bonus(int vocab) const;
  void step_bonus(const Walk& w, int vocab, std::vector<std::pair<int32_t, float>>& out) const;

 private:
  struct Node {
    std::unordered_map<int32_t, int32_t> children;
    int depth = 0;
  };
  std::vector<Node> nodes_{Node{}};
  size_t sequences_ = 0;
  float boost_ = 2.0f;
};

// Replace structurally invalid UTF-8 by '?' (core/transcriber.cpp:1489-1541).
std::string sanitize_utf8(const std::string& text);

}  // namespace msb
