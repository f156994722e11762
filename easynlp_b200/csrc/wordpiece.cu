// Native WordPiece tokenizer (host code; SURVEY.md 8f.2: the step before the hot path).  Restates BertTokenizer of the reference
// (easynlp/modelzoo/models/bert/tokenization_bert.py:67-504: BasicTokenizer -- clean, CJK spacing, lower-case + accent stripping,
// punctuation split -- and WordpieceTokenizer, greedy longest match first) for the call the CLIP application makes
// (appzoo/clip/data.py:262-264: padding='max_length', truncation=True): [CLS] ids [SEP] + padding, attention mask.
// The per-code-point predicates come from a table generated with the same Python unicodedata the reference evaluates
// (tools/gen_unicode_table.py), so the result is exact on the BMP; CJK extension planes are handled by rule, any other supplementary
// code point (and capital sigma, whose lower-casing is context sensitive) makes the call return CLIPK_WP_FALLBACK and the Python
// tokenizer takes that text.  Batches are encoded on several host threads.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "common.cuh"
#include "../../include/clipk.h"

namespace {

struct UniEntry { uint8_t flags, mlen; uint32_t moff; };

struct WordPiece {
  std::unordered_map<std::string, int> vocab;
  std::vector<UniEntry> uni;
  std::vector<uint32_t> pool;
  int lower = 1;
  int unk = 0, cls = 0, sep = 0, pad = 0;
  std::string unk_s = "[UNK]", cls_s = "[CLS]", sep_s = "[SEP]", pad_s = "[PAD]", mask_s = "[MASK]";
};

bool is_cjk(uint32_t cp) {
  return (cp >= 0x4E00 && cp <= 0x9FFF) || (cp >= 0x3400 && cp <= 0x4DBF) || (cp >= 0x20000 && cp <= 0x2A6DF) || (cp >= 0x2A700 && cp <= 0x2B73F) ||
         (cp >= 0x2B740 && cp <= 0x2B81F) || (cp >= 0x2B820 && cp <= 0x2CEAF) || (cp >= 0xF900 && cp <= 0xFAFF) || (cp >= 0x2F800 && cp <= 0x2FA1F);
}

bool decode_utf8(const char* s, std::vector<uint32_t>& out) {
  const unsigned char* p = (const unsigned char*)s;
  while (*p) {
    uint32_t cp; int n;
    if (*p < 0x80) { cp = *p; n = 1; }
    else if ((*p >> 5) == 6) { cp = *p & 31; n = 2; }
    else if ((*p >> 4) == 14) { cp = *p & 15; n = 3; }
    else if ((*p >> 3) == 30) { cp = *p & 7; n = 4; }
    else return false;
    for (int i = 1; i < n; ++i) { if ((p[i] >> 6) != 2) return false; cp = (cp << 6) | (p[i] & 63); }
    out.push_back(cp);
    p += n;
  }
  return true;
}
void append_utf8(std::string& s, uint32_t cp) {
  if (cp < 0x80) s.push_back((char)cp);
  else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 63))); }
  else if (cp < 0x10000) { s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
  else { s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 63))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
}
std::string to_utf8(const uint32_t* b, const uint32_t* e) { std::string s; for (; b < e; ++b) append_utf8(s, *b); return s; }

// flags of a code point; supplementary planes: CJK by rule, everything else unsupported (16)
inline uint8_t flags_of(const WordPiece& w, uint32_t cp) { return cp < 65536 ? w.uni[cp].flags : (is_cjk(cp) ? 0 : 16); }

// -> number of tokens written (without specials), or CLIPK_WP_FALLBACK
int tokenize(const WordPiece& w, const char* text, std::vector<int>& ids) {
  std::vector<uint32_t> cps;
  if (!decode_utf8(text, cps)) return CLIPK_WP_FALLBACK;
  // clean + CJK spacing (tokenization_bert.py: _clean_text, _tokenize_chinese_chars)
  std::vector<uint32_t> t; t.reserve(cps.size() * 3);
  for (uint32_t cp : cps) {
    const uint8_t f = flags_of(w, cp);
    if (f & 16) return CLIPK_WP_FALLBACK;
    if (f & 2) continue;
    if (f & 1) { t.push_back(' '); continue; }
    if (is_cjk(cp)) { t.push_back(' '); t.push_back(cp); t.push_back(' '); }
    else t.push_back(cp);
  }
  // whitespace split (str.split(): any code point with str.isspace)
  std::vector<std::vector<uint32_t>> words;
  {
    std::vector<uint32_t> cur;
    for (uint32_t cp : t) {
      const bool sp = cp < 65536 ? (w.uni[cp].flags & 8) != 0 : false;
      if (sp) { if (!cur.empty()) { words.push_back(cur); cur.clear(); } }
      else cur.push_back(cp);
    }
    if (!cur.empty()) words.push_back(cur);
  }
  ids.clear();
  std::vector<std::vector<uint32_t>> toks;
  for (auto& word : words) {
    const std::string ws = to_utf8(word.data(), word.data() + word.size());
    if (ws == w.unk_s || ws == w.sep_s || ws == w.pad_s || ws == w.cls_s || ws == w.mask_s) { toks.push_back(word); continue; }   // never_split
    std::vector<uint32_t> lw;
    if (w.lower) {
      for (uint32_t cp : word) {
        if (cp < 65536 && w.uni[cp].mlen != 255) { const UniEntry& e = w.uni[cp]; for (int i = 0; i < e.mlen; ++i) lw.push_back(w.pool[e.moff + i]); }
        else lw.push_back(cp);
      }
    } else lw = word;
    // punctuation split; the pieces are re-split on whitespace like `" ".join(tokens).strip().split()` does
    std::vector<uint32_t> cur;
    auto flush = [&]() { if (!cur.empty()) { toks.push_back(cur); cur.clear(); } };
    for (uint32_t cp : lw) {
      const uint8_t f = cp < 65536 ? w.uni[cp].flags : 0;
      if (f & 8) { flush(); continue; }
      if (f & 4) { flush(); toks.push_back(std::vector<uint32_t>{cp}); }
      else cur.push_back(cp);
    }
    flush();
  }
  // WordPiece: greedy longest match first, "##" continuation prefix, > 100 characters -> [UNK]
  for (auto& tok : toks) {
    const std::string ts = to_utf8(tok.data(), tok.data() + tok.size());
    if (ts == w.unk_s || ts == w.sep_s || ts == w.pad_s || ts == w.cls_s || ts == w.mask_s) {
      auto it = w.vocab.find(ts); ids.push_back(it == w.vocab.end() ? w.unk : it->second); continue;
    }
    if (tok.size() > 100) { ids.push_back(w.unk); continue; }
    std::vector<int> sub;
    size_t start = 0; bool bad = false;
    while (start < tok.size()) {
      size_t end = tok.size(); int found = -1;
      while (start < end) {
        std::string s = start > 0 ? "##" : "";
        for (size_t i = start; i < end; ++i) append_utf8(s, tok[i]);
        auto it = w.vocab.find(s);
        if (it != w.vocab.end()) { found = it->second; break; }
        --end;
      }
      if (found < 0) { bad = true; break; }
      sub.push_back(found);
      start = end;
    }
    if (bad) ids.push_back(w.unk); else ids.insert(ids.end(), sub.begin(), sub.end());
  }
  return (int)ids.size();
}

int encode_one(const WordPiece& w, const char* text, int max_length, long long* ids_out, long long* mask_out) {
  std::vector<int> ids;
  const int n = tokenize(w, text, ids);
  if (n < 0) return n;
  int keep = n > max_length - 2 ? max_length - 2 : n;       // truncation=True
  if (keep < 0) keep = 0;
  int k = 0;
  if (max_length >= 1) { ids_out[k] = w.cls; mask_out[k] = 1; ++k; }
  for (int i = 0; i < keep && k < max_length; ++i, ++k) { ids_out[k] = ids[i]; mask_out[k] = 1; }
  if (k < max_length) { ids_out[k] = w.sep; mask_out[k] = 1; ++k; }
  const int used = k;
  for (; k < max_length; ++k) { ids_out[k] = w.pad; mask_out[k] = 0; }
  return used;
}

}  // namespace

extern "C" void* clipk_wp_create(const char* vocab_path, const char* unicode_table_path, int do_lower_case) {
  WordPiece* w = new WordPiece();
  w->lower = do_lower_case;
  FILE* f = fopen(vocab_path, "rb");
  if (!f) { clipk::set_error("wp_create: cannot open vocab %s", vocab_path); delete w; return nullptr; }
  {
    std::string line; int c, idx = 0;
    while ((c = fgetc(f)) != EOF) {
      if (c == '\n') { w->vocab.emplace(line, idx++); line.clear(); }      // like Python: token.rstrip("\n"); a later duplicate overwrites
      else line.push_back((char)c);
    }
    if (!line.empty()) w->vocab.emplace(line, idx++);
  }
  fclose(f);
  // Python's OrderedDict assignment keeps the LAST index of a duplicated token; unordered_map::emplace keeps the first -> redo with overwrite
  {
    f = fopen(vocab_path, "rb");
    std::string line; int c, idx = 0;
    while ((c = fgetc(f)) != EOF) { if (c == '\n') { w->vocab[line] = idx++; line.clear(); } else line.push_back((char)c); }
    if (!line.empty()) w->vocab[line] = idx++;
    fclose(f);
  }
  f = fopen(unicode_table_path, "rb");
  if (!f) { clipk::set_error("wp_create: cannot open unicode table %s", unicode_table_path); delete w; return nullptr; }
  char magic[8]; uint32_t pool_len = 0;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "CLPKUNI1", 8) != 0 || fread(&pool_len, 4, 1, f) != 1) { fclose(f); clipk::set_error("wp_create: bad unicode table"); delete w; return nullptr; }
  w->uni.resize(65536);
  for (int cp = 0; cp < 65536; ++cp) {
    unsigned char rec[6];
    if (fread(rec, 1, 6, f) != 6) { fclose(f); clipk::set_error("wp_create: truncated unicode table"); delete w; return nullptr; }
    w->uni[cp].flags = rec[0]; w->uni[cp].mlen = rec[1]; memcpy(&w->uni[cp].moff, rec + 2, 4);
  }
  w->pool.resize(pool_len);
  if (pool_len && fread(w->pool.data(), 4, pool_len, f) != pool_len) { fclose(f); clipk::set_error("wp_create: truncated unicode pool"); delete w; return nullptr; }
  fclose(f);
  auto id = [&](const std::string& s, int dflt) { auto it = w->vocab.find(s); return it == w->vocab.end() ? dflt : it->second; };
  w->unk = id(w->unk_s, 0); w->pad = id(w->pad_s, 0);
  if (!w->vocab.count(w->cls_s) || !w->vocab.count(w->sep_s)) { clipk::set_error("wp_create: vocab lacks [CLS] / [SEP]"); delete w; return nullptr; }
  w->cls = w->vocab[w->cls_s]; w->sep = w->vocab[w->sep_s];
  return w;
}

extern "C" void clipk_wp_destroy(void* h) { delete (WordPiece*)h; }

extern "C" int clipk_wp_encode(void* h, const char* utf8_text, int max_length, long long* input_ids, long long* attention_mask) {
  if (!h || !utf8_text || max_length < 2) { clipk::set_error("wp_encode: bad arguments"); return CLIPK_ERR_ARG; }
  return encode_one(*(WordPiece*)h, utf8_text, max_length, input_ids, attention_mask);
}

extern "C" int clipk_wp_encode_batch(void* h, const char* const* texts, int n, int max_length, long long* input_ids, long long* attention_mask, int* status,
                                     int threads) {
  if (!h || !texts || n < 0 || max_length < 2) { clipk::set_error("wp_encode_batch: bad arguments"); return CLIPK_ERR_ARG; }
  const WordPiece& w = *(WordPiece*)h;
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  auto work = [&](int t) {
    for (int i = t; i < n; i += threads)
      status[i] = encode_one(w, texts[i], max_length, input_ids + (long long)i * max_length, attention_mask + (long long)i * max_length);
  };
  if (threads == 1) { work(0); return 0; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
  for (auto& th : pool) th.join();
  return 0;
}
