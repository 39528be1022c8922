#include "tokenizer.h"

#include <algorithm>
#include <cctype>
#include <cmath>
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
  build_encoder_index();
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

void Tokenizer::build_encoder_index() {
  by_first_byte_.assign(256, {});
  for (size_t i = 0; i < pieces_.size(); i++)
    if (!pieces_[i].empty()) by_first_byte_[(uint8_t)pieces_[i][0]].push_back((int32_t)i);
  byte_base_ = -1;
  for (size_t start = 0; start + 256 <= pieces_.size() && byte_base_ < 0; start++) {
    bool whole = true;
    for (size_t o = 0; o < 256 && whole; o++)
      whole = pieces_[start + o].size() == 1 && (uint8_t)pieces_[start + o][0] == o;
    if (whole) byte_base_ = (int32_t)start;
  }
  merge_ids_.clear();
  if (byte_base_ >= 0)
    for (size_t i = (size_t)byte_base_ + 256; i < pieces_.size(); i++)
      if (!pieces_[i].empty()) merge_ids_.emplace(pieces_[i], (int32_t)i);  // first (lowest) id wins
}

static std::string spaces_to_marker(const std::string& text) {
  std::string out;
  for (char c : text) {
    if (c == ' ') out += "\xE2\x96\x81";
    else out.push_back(c);
  }
  return out;
}

std::vector<int32_t> Tokenizer::encode_longest_match(const std::string& text) const {
  const std::string t = spaces_to_marker(text);
  std::vector<int32_t> out;
  size_t pos = 0;
  while (pos < t.size()) {
    size_t best_len = 0;
    int32_t best = -1;
    for (int32_t id : by_first_byte_[(uint8_t)t[pos]]) {
      const std::string& p = pieces_[(size_t)id];
      if (p.size() > best_len && p.size() <= t.size() - pos && t.compare(pos, p.size(), p) == 0) {
        best_len = p.size();
        best = id;
      }
    }
    if (best < 0) throw std::runtime_error("No match found for remaining bytes " + t.substr(pos));
    out.push_back(best);
    pos += best_len;
  }
  return out;
}

std::vector<int32_t> Tokenizer::text_to_tokens(const std::string& text, bool bpe) const {
  if (!bpe || byte_base_ < 0) return encode_longest_match(text);
  const std::string t = spaces_to_marker(text);
  // one piece per UTF-8 character (a byte that cannot start a sequence stands alone)
  std::vector<std::string> parts;
  for (size_t o = 0; o < t.size();) {
    const uint8_t lead = (uint8_t)t[o];
    size_t n = (lead & 0x80) == 0 ? 1 : (lead & 0xE0) == 0xC0 ? 2 : (lead & 0xF0) == 0xE0 ? 3 : (lead & 0xF8) == 0xF0 ? 4 : 1;
    n = std::min(n, t.size() - o);
    parts.push_back(t.substr(o, n));
    o += n;
  }
  // replay the merges: always join the adjacent pair whose spelling has the lowest id (leftmost on ties)
  while (parts.size() > 1) {
    int32_t best_id = -1;
    size_t best_pos = 0;
    for (size_t i = 0; i + 1 < parts.size(); i++) {
      const auto it = merge_ids_.find(parts[i] + parts[i + 1]);
      if (it != merge_ids_.end() && (best_id < 0 || it->second < best_id)) {
        best_id = it->second;
        best_pos = i;
      }
    }
    if (best_id < 0) break;
    parts[best_pos] += parts[best_pos + 1];
    parts.erase(parts.begin() + (long)best_pos + 1);
  }
  std::vector<int32_t> out;
  for (const std::string& part : parts) {
    const auto it = merge_ids_.find(part);
    if (it != merge_ids_.end()) {
      out.push_back(it->second);
    } else {
      for (char c : part) out.push_back(byte_base_ + (uint8_t)c);
    }
  }
  return out;
}

// ---- key terms from a passage ----
namespace {
bool word_byte(unsigned char b) { return b >= 0x80 || std::isalpha(b) != 0; }
bool digit_byte(unsigned char b) { return std::isdigit(b) != 0; }
bool joiner_byte(unsigned char b) { return b == '\'' || b == '-'; }
std::string strip_joiners(const std::string& w) {
  size_t b = 0, e = w.size();
  while (b < e && joiner_byte((unsigned char)w[b])) b++;
  while (e > b && joiner_byte((unsigned char)w[e - 1])) e--;
  return w.substr(b, e - b);
}
std::string drop_possessive(const std::string& w) {
  const size_t n = w.size();
  if (n >= 2 && w[n - 2] == '\'' && (w[n - 1] == 's' || w[n - 1] == 'S')) return w.substr(0, n - 2);
  if (n >= 1 && w[n - 1] == '\'') return w.substr(0, n - 1);
  return w;
}
void replace_every(std::string& s, const std::string& from, const std::string& to) {
  size_t pos = 0;
  while ((pos = s.find(from, pos)) != std::string::npos) {
    s.replace(pos, from.size(), to);
    pos += to.size();
  }
}
std::vector<std::string> passage_words(const std::string& text) {
  std::string t = text;
  // typographic punctuation that arrives as multi-byte UTF-8 in prose
  static const char* const kFold[][2] = {
      {"\xe2\x80\x99", "'"}, {"\xe2\x80\x98", "'"}, {"\xe2\x80\x9c", " "}, {"\xe2\x80\x9d", " "},
      {"\xe2\x80\x93", " "}, {"\xe2\x80\x94", " "}, {"\xe2\x80\xa6", " "}, {"\xc2\xa0", " "}};
  for (const auto& m : kFold) replace_every(t, m[0], m[1]);
  std::vector<std::string> words;
  std::string cur;
  auto flush = [&]() {
    if (cur.empty()) return;
    const std::string w = strip_joiners(drop_possessive(strip_joiners(cur)));
    cur.clear();
    size_t chars = 0;
    bool has_digit = false;
    for (char ch : w) {
      if (((unsigned char)ch & 0xc0) != 0x80) chars++;
      has_digit = has_digit || digit_byte((unsigned char)ch);
    }
    if (chars < 3 || has_digit) return;
    words.push_back(w);
  };
  for (char ch : t) {
    const unsigned char b = (unsigned char)ch;
    if (word_byte(b) || digit_byte(b) || (joiner_byte(b) && !cur.empty())) cur.push_back(ch);
    else flush();
  }
  flush();
  return words;
}
}  // namespace

std::vector<std::string> extract_key_terms(const std::string& context, int32_t max_terms, const Tokenizer& tokenizer) {
  const size_t limit = (size_t)(max_terms > 0 ? max_terms : 200);
  struct Form { size_t count = 0, first = 0; };
  struct Group { std::string term; size_t term_count = 0, occurrences = 0, subwords = 0, first = 0; bool seen = false; };
  const std::vector<std::string> words = passage_words(context);
  std::unordered_map<std::string, Form> forms;
  for (size_t i = 0; i < words.size(); i++) {
    auto it = forms.emplace(words[i], Form{0, i}).first;
    it->second.count++;
  }
  std::unordered_map<std::string, Group> groups;
  for (const auto& kv : forms) {
    std::string folded = kv.first;
    for (char& ch : folded)
      if ((unsigned char)ch < 0x80) ch = (char)std::tolower((unsigned char)ch);
    Group& g = groups[folded];
    g.occurrences += kv.second.count;
    if (!g.seen || kv.second.count > g.term_count || (kv.second.count == g.term_count && kv.second.first < g.first)) {
      g.term = kv.first;
      g.term_count = kv.second.count;
      g.first = kv.second.first;
      g.seen = true;
    }
  }
  std::vector<Group> picked;
  for (auto& kv : groups) {
    Group& g = kv.second;
    try {
      g.subwords = tokenizer.text_to_tokens(" " + g.term, /*bpe=*/true).size();
    } catch (const std::exception&) {
      g.subwords = 0;  // an unspellable word costs that word and nothing else
    }
    if (g.subwords >= 2) picked.push_back(g);
  }
  std::sort(picked.begin(), picked.end(), [](const Group& a, const Group& b) {
    if (a.occurrences != b.occurrences) return a.occurrences > b.occurrences;
    if (a.subwords != b.subwords) return a.subwords > b.subwords;
    return a.first < b.first;
  });
  std::vector<std::string> terms;
  for (const Group& g : picked) {
    if (terms.size() >= limit) break;
    terms.push_back(g.term);
  }
  return terms;
}

// ---- KeytermBiaser ----
void KeytermBiaser::clear() {
  nodes_.assign(1, Node{});
  sequences_ = 0;
}

void KeytermBiaser::add_token_sequence(const std::vector<int32_t>& tokens) {
  if (tokens.empty()) return;
  int32_t at = 0;
  for (int32_t tok : tokens) {
    auto it = nodes_[(size_t)at].children.find(tok);
    if (it != nodes_[(size_t)at].children.end()) {
      at = it->second;
      continue;
    }
    const int depth = nodes_[(size_t)at].depth + 1;
    const int32_t child = (int32_t)nodes_.size();
    nodes_[(size_t)at].children.emplace(tok, child);
    nodes_.push_back(Node{});
    nodes_.back().depth = depth;
    at = child;
  }
  sequences_++;
}

std::vector<std::string> KeytermBiaser::variants_for_term(const std::string& term) {
  const size_t b = term.find_first_not_of(" \t");
  if (b == std::string::npos) return {};
  const std::string t = term.substr(b, term.find_last_not_of(" \t") - b + 1);
  if (t.compare(0, 3, "\xE2\x96\x81") == 0) return {t};  // already anchored to a word start
  return {t, " " + t};
}

void KeytermBiaser::apply(const Walk& w, float* logits, int vocab) const {
  if (logits == nullptr || sequences_ == 0) return;
  std::vector<std::pair<int32_t, float>> pending;
  for (int32_t node_index : w.active) {
    const Node& node = nodes_[(size_t)node_index];
    const int depth = node.depth + 1;
    const float bonus = boost_ * (1.0f + std::log((float)depth));
    for (const auto& child : node.children) {
      const int32_t tok = child.first;
      if (tok < 0 || tok >= vocab) continue;
      bool merged = false;
      for (auto& pb : pending) {
        if (pb.first == tok) {
          pb.second = std::max(pb.second, bonus);
          merged = true;
          break;
        }
      }
      if (!merged) pending.emplace_back(tok, bonus);
    }
  }
  for (const auto& pb : pending) logits[pb.first] += pb.second;
}

std::vector<float> KeytermBiaser::root_bonus(int vocab) const {
  std::vector<float> out((size_t)vocab, 0.f);
  if (sequences_ == 0) return out;
  const float bonus = boost_ * (1.0f + std::log(1.0f));  // depth 1
  for (const auto& child : nodes_[0].children)
    if (child.first >= 0 && child.first < vocab) out[(size_t)child.first] = bonus;
  return out;
}

void KeytermBiaser::step_bonus(const Walk& w, int vocab, std::vector<std::pair<int32_t, float>>& out) const {
  out.clear();
  if (sequences_ == 0) return;
  for (int32_t node_index : w.active) {
    if (node_index == 0) continue;  // the root's children are the shared part
    const Node& node = nodes_[(size_t)node_index];
    const float bonus = boost_ * (1.0f + std::log((float)(node.depth + 1)));
    for (const auto& child : node.children) {
      const int32_t tok = child.first;
      if (tok < 0 || tok >= vocab) continue;
      bool merged = false;
      for (auto& pb : out) {
        if (pb.first == tok) {
          pb.second = std::max(pb.second, bonus);
          merged = true;
          break;
        }
      }
      if (!merged) out.emplace_back(tok, bonus);
    }
  }
  // merged bonus -> what it adds on top of the root's share of the same token (apply() keeps the largest)
  const float root = boost_ * (1.0f + std::log(1.0f));
  size_t keep = 0;
  for (auto& pb : out) {
    const bool root_child = nodes_[0].children.find(pb.first) != nodes_[0].children.end();
    const float delta = root_child ? std::max(pb.second, root) - root : pb.second;
    if (delta != 0.0f) out[keep++] = {pb.first, delta};
  }
  out.resize(keep);
}

void KeytermBiaser::advance(Walk& w, int32_t token) const {
  if (sequences_ == 0) return;
  std::vector<int32_t> next{0};  // the root stays active: a term can start at any token
  for (int32_t node_index : w.active) {
    const auto it = nodes_[(size_t)node_index].children.find(token);
    if (it != nodes_[(size_t)node_index].children.end()) next.push_back(it->second);
  }
  w.active.swap(next);
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
