// nibtools.h -- 4-bit .nib reader with the public interface of the reference's class nib
// (reference src/nibtools.h:12-118): open() status codes index the same errormsg[] table,
// getBase() takes a 0-based position, bases decode to upper case (soft-mask bit ignored).
// Written from scratch: the file is memory-loaded once (the reference seeks and reads one byte per
// base, src/nibtools.cc:38-64) so the packed payload can be handed to the GPU (bkid_set_nib).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#define NIB_MAGIC 0x6be93d3aUL

extern const std::string nib_errormsg[5];   // "", "wrong format", "cannot open file", "file is not open", "position beyond sequence boundary"

class nib {
 public:
  int open(std::string filename);               // 0 ok, 1 wrong format, 2 cannot open
  void close() { data_.clear(); open_ = false; cursor_ = 0; }
  int getBase(char *base, unsigned long pos);   // 0 ok, 3 not open, 4 beyond boundary, 1 bad code ('N' written)
  int nextBase(char *base);
  unsigned long size() { return open_ ? n_bases_ : 0; }
  // extensions used by the GPU path
  const uint8_t *payload() const { return data_.data(); }   // packed bases, high nibble first
  size_t payload_bytes() const { return data_.size(); }

 private:
  std::vector<uint8_t> data_;
  unsigned long n_bases_ = 0, cursor_ = 0;
  bool open_ = false;
  static int decode(char *out, int code);
};
