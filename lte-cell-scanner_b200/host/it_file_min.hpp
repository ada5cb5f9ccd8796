// it_file_min.hpp - reader/writer for the IT++ it_file v3 records CellSearch uses:
// capbuf_NNNN.it = { "capbuf": dcvec, "fc": ivec }  (reference src/capbuf.cpp:98-115,187-197).
// Format: "IT++" u8(3) then per variable {u64 hdr, u64 data, u64 block,
// name\0 type\0 desc\0, payload}; vectors carry a u64 length.  Little-endian.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "itpp_min.hpp"

namespace lcs_it {

inline bool read_all(const std::string& path, std::vector<unsigned char>& buf) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? n : 0);
  size_t got = n > 0 ? std::fread(buf.data(), 1, n, f) : 0;
  std::fclose(f);
  return got == (size_t)n;
}

// Finds variable `name`; returns payload pointer/size and the IT++ type string.
inline bool find_var(const std::vector<unsigned char>& b, const std::string& name, std::string& type,
                     const unsigned char*& payload, uint64_t& payload_bytes) {
  if (b.size() < 5 || std::memcmp(b.data(), "IT++", 4) != 0 || b[4] != 3) return false;
  size_t pos = 5;
  while (pos + 24 <= b.size()) {
    uint64_t hdr, data, block;
    std::memcpy(&hdr, &b[pos], 8);
    std::memcpy(&data, &b[pos + 8], 8);
    std::memcpy(&block, &b[pos + 16], 8);
    if (block == 0 || pos + block > b.size()) return false;
    const char* s = reinterpret_cast<const char*>(&b[pos + 24]);
    std::string nm(s);
    std::string ty(s + nm.size() + 1);
    if (nm == name) {
      type = ty;
      payload = &b[pos + hdr];
      payload_bytes = data;
      return true;
    }
    pos += block;
  }
  return false;
}

inline bool read_capbuf(const std::string& path, itpp::cvec& capbuf, int& fc) {
  std::vector<unsigned char> b;
  if (!read_all(path, b)) return false;
  std::string ty;
  const unsigned char* p;
  uint64_t nbytes;
  if (!find_var(b, "capbuf", ty, p, nbytes) || ty != "dcvec") return false;
  uint64_t n;
  std::memcpy(&n, p, 8);
  if (8 + n * 16 > nbytes) return false;
  capbuf.set_size((int)n);
  std::memcpy(capbuf._data(), p + 8, n * 16);
  if (!find_var(b, "fc", ty, p, nbytes) || ty != "ivec") return false;
  std::memcpy(&n, p, 8);
  if (n < 1) return false;
  int32_t v;
  std::memcpy(&v, p + 8, 4);
  fc = v;
  return true;
}

inline void put_var(std::vector<unsigned char>& out, const std::string& name, const std::string& type,
                    const void* payload, uint64_t n_items, size_t item_bytes) {
  std::string strs = name + '\0' + type + '\0' + '\0';
  uint64_t hdr = 24 + strs.size(), data = 8 + n_items * item_bytes, block = hdr + data;
  size_t o = out.size();
  out.resize(o + block);
  std::memcpy(&out[o], &hdr, 8);
  std::memcpy(&out[o + 8], &data, 8);
  std::memcpy(&out[o + 16], &block, 8);
  std::memcpy(&out[o + 24], strs.data(), strs.size());
  std::memcpy(&out[o + hdr], &n_items, 8);
  std::memcpy(&out[o + hdr + 8], payload, n_items * item_bytes);
}

inline bool write_capbuf(const std::string& path, const itpp::cvec& capbuf, int fc) {   // capbuf.cpp:187-197
  std::vector<unsigned char> out = {'I', 'T', '+', '+', 3};
  put_var(out, "capbuf", "dcvec", capbuf._data(), (uint64_t)capbuf.length(), 16);
  int32_t v = fc;
  put_var(out, "fc", "ivec", &v, 1, 4);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}

}  // namespace lcs_it
