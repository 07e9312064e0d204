#include "nibtools.h"

#include <cstdio>

const std::string nib_errormsg[5] = {"", "wrong format", "cannot open file", "file is not open", "position beyond sequence boundary"};

int nib::decode(char *out, int code)
{
  // T=0 C=1 A=2 G=3 N=4, +8 = soft-masked (reference src/nibtools.h:23-59)
  switch (code & 0xff) {
    case 0: case 8: *out = 'T'; return 0;
    case 1: case 9: *out = 'C'; return 0;
    case 2: case 10: *out = 'A'; return 0;
    case 3: case 11: *out = 'G'; return 0;
    case 4: *out = 'N'; return 0;
    default: *out = 'N'; return 1;
  }
}

int nib::open(std::string filename)
{
  close();
  FILE *f = fopen(filename.c_str(), "rb");
  if (!f) return 2;
  unsigned char h[8];
  if (fread(h, 1, 8, f) != 8) { fclose(f); return 1; }
  unsigned long magic = h[0] | (h[1] << 8) | (h[2] << 16) | ((unsigned long)h[3] << 24);
  if (magic != NIB_MAGIC) { fclose(f); return 1; }
  n_bases_ = h[4] | (h[5] << 8) | (h[6] << 16) | ((unsigned long)h[7] << 24);
  fseek(f, 0, SEEK_END);
  unsigned long fsz = (unsigned long)ftell(f);
  unsigned long i = fsz * 2 - 16;                      // reference src/nibtools.cc:30-33
  if (!(n_bases_ == i || n_bases_ + 1 == i)) { fclose(f); return 1; }
  fseek(f, 8, SEEK_SET);
  data_.resize(fsz - 8);
  size_t got = fread(data_.data(), 1, data_.size(), f);
  fclose(f);
  if (got != data_.size()) { data_.clear(); return 1; }
  open_ = true;
  cursor_ = 0;
  return 0;
}

int nib::getBase(char *base, unsigned long pos)
{
  if (!open_) return 3;
  if (pos >= n_bases_) return 4;
  int b = data_[pos / 2];
  return decode(base, (pos % 2 == 0) ? (b >> 4) : (b & 0x0f));
}

int nib::nextBase(char *base)
{
  if (!open_) return 3;
  int r = getBase(base, cursor_);
  ++cursor_;
  return r;
}
