// c2ray_io.h -- Fortran `form="unformatted"` sequential records, host side only: the on-disk format of the reference's
// iteration dumps (files_for_3D/evolve.F90:233-275 write_iteration_dump, :279-367 start_from_dump) and of its output
// streams 2 and 3 (files_for_3D/output.F90:249-379), so that a Fortran host can resume from / inspect what the GPU
// path wrote and vice versa.
//
// Record layout (gfortran and ifort defaults): int32 byte count, payload, int32 byte count.  A record longer than
// 2^31-9 bytes is split into subrecords; the leading marker of a subrecord is negative when another subrecord
// follows, the trailing marker is negative when one preceded it.  (xh_av of a 512^3 mesh is 2.1 GB: this matters.)
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace c2io {

constexpr uint64_t MAX_SUBRECORD = 2147483639ull;  // 2^31 - 9

class RecordWriter {
 public:
  explicit RecordWriter(uint64_t max_sub = MAX_SUBRECORD) : max_sub_(max_sub ? max_sub : MAX_SUBRECORD) {}
  ~RecordWriter() { close(); }
  bool open(const std::string& path) {
    f_ = fopen(path.c_str(), "wb");
    return f_ != nullptr;
  }
  bool close() {
    bool ok = true;
    if (f_) { ok = fclose(f_) == 0 && !in_record_; f_ = nullptr; }
    return ok;
  }
  // a record is begun with its total size, filled by any number of put() calls, and closes itself when full
  bool begin(uint64_t total) {
    if (!f_ || in_record_) return false;
    total_ = total; done_ = 0; first_sub_ = true; in_record_ = true;
    if (!start_sub()) return false;
    if (total == 0) return end_sub();
    return true;
  }
  bool put(const void* data, uint64_t n) {
    const char* p = static_cast<const char*>(data);
    if (!in_record_ || done_ + n > total_) return false;
    while (n) {
      const uint64_t k = n < sub_len_ - sub_done_ ? n : sub_len_ - sub_done_;
      if (fwrite(p, 1, k, f_) != k) return false;
      p += k; n -= k; sub_done_ += k; done_ += k;
      if (sub_done_ == sub_len_ && !end_sub()) return false;
    }
    return true;
  }
  bool record(const void* data, uint64_t n) { return begin(n) && (n == 0 || put(data, n)); }
  bool in_record() const { return in_record_; }

 private:
  bool start_sub() {
    const uint64_t rem = total_ - done_;
    sub_len_ = rem < max_sub_ ? rem : max_sub_;
    sub_done_ = 0;
    const int32_t m = rem > max_sub_ ? -(int32_t)sub_len_ : (int32_t)sub_len_;
    return fwrite(&m, 4, 1, f_) == 1;
  }
  bool end_sub() {
    const int32_t m = first_sub_ ? (int32_t)sub_len_ : -(int32_t)sub_len_;
    if (fwrite(&m, 4, 1, f_) != 1) return false;
    first_sub_ = false;
    if (done_ < total_) return start_sub();
    in_record_ = false;
    return true;
  }
  FILE* f_ = nullptr;
  uint64_t max_sub_, total_ = 0, done_ = 0, sub_len_ = 0, sub_done_ = 0;
  bool first_sub_ = true, in_record_ = false;
};

class RecordReader {
 public:
  ~RecordReader() { close(); }
  bool open(const std::string& path) {
    f_ = fopen(path.c_str(), "rb");
    return f_ != nullptr;
  }
  void close() { if (f_) { fclose(f_); f_ = nullptr; } }
  // begin a record that must hold exactly `total` bytes; get() consumes it piecewise
  bool begin(uint64_t total) {
    if (!f_ || in_record_) return false;
    total_ = total; done_ = 0; first_sub_ = true; in_record_ = true;
    if (!start_sub()) return false;
    if (total == 0) return end_sub();
    return true;
  }
  bool get(void* data, uint64_t n) {
    char* p = static_cast<char*>(data);
    if (!in_record_ || done_ + n > total_) return false;
    while (n) {
      const uint64_t k = n < sub_len_ - sub_done_ ? n : sub_len_ - sub_done_;
      if (k == 0) return false;  // the file's record is shorter than expected
      if (fread(p, 1, k, f_) != k) return false;
      p += k; n -= k; sub_done_ += k; done_ += k;
      if (sub_done_ == sub_len_ && !end_sub()) return false;
    }
    return true;
  }
  bool record(void* data, uint64_t n) { return begin(n) && (n == 0 || get(data, n)); }

 private:
  bool start_sub() {
    int32_t m;
    if (fread(&m, 4, 1, f_) != 1) return false;
    more_ = m < 0;
    sub_len_ = (uint64_t)(m < 0 ? -(int64_t)m : (int64_t)m);
    sub_done_ = 0;
    if (sub_len_ > total_ - done_) return false;              // record longer than expected
    if (!more_ && sub_len_ != total_ - done_) return false;   // record shorter than expected
    return true;
  }
  bool end_sub() {
    int32_t m;
    if (fread(&m, 4, 1, f_) != 1) return false;
    const uint64_t len = (uint64_t)(m < 0 ? -(int64_t)m : (int64_t)m);
    if (len != sub_len_ || (m < 0) == first_sub_) {
      if (!(len == sub_len_ && sub_len_ == 0)) return false;  // corrupt markers
    }
    first_sub_ = false;
    if (more_) return start_sub();
    if (done_ != total_) return false;
    in_record_ = false;
    return true;
  }
  FILE* f_ = nullptr;
  uint64_t total_ = 0, done_ = 0, sub_len_ = 0, sub_done_ = 0;
  bool first_sub_ = true, in_record_ = false, more_ = false;
};

// write(file1,"(f6.3)") zred_now ; trim(adjustl(file1))  -- output.F90:264.  A value that does not fit six characters
// prints as asterisks in Fortran.
inline std::string f6_3(double z) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.3f", z);
  if (strlen(buf) > 6) return "******";
  return buf;
}

}  // namespace c2io
