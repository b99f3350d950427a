// Minimal stand-in for boost::iostreams::filtering_istream + gzip_decompressor, on zlib.
// Only the usage pattern of src/preselect.cpp:761-766 is supported:
//   filtering_istream in; in.push(gzip_decompressor()); in.push(std::ifstream&);
// Test infrastructure only.
#ifndef HLM_SHIM_BOOST_FILTERING_STREAM_HPP
#define HLM_SHIM_BOOST_FILTERING_STREAM_HPP
#include <zlib.h>

#include <istream>
#include <stdexcept>
#include <streambuf>
#include <vector>
namespace boost {
namespace iostreams {
struct gzip_decompressor {};
struct gzip_error : public std::runtime_error {
  gzip_error() : std::runtime_error("gzip error") {}
};
class inflate_buf : public std::streambuf {
 public:
  inflate_buf() : src_(nullptr), in_(1 << 16), out_(1 << 16), ok_(false), eof_(false) {}
  ~inflate_buf() {
    if (ok_) inflateEnd(&z_);
  }
  void attach(std::istream& s) {
    src_ = &s;
    z_ = z_stream();
    ok_ = inflateInit2(&z_, 15 + 32) == Z_OK;
    if (!ok_) throw gzip_error();
  }

 protected:
  int_type underflow() override {
    if (!src_ || eof_) return traits_type::eof();
    for (;;) {
      if (z_.avail_in == 0) {
        src_->read(in_.data(), (std::streamsize)in_.size());
        z_.avail_in = (uInt)src_->gcount();
        z_.next_in = (Bytef*)in_.data();
      }
      z_.avail_out = (uInt)out_.size();
      z_.next_out = (Bytef*)out_.data();
      int rc = inflate(&z_, Z_NO_FLUSH);
      size_t got = out_.size() - z_.avail_out;
      if (rc == Z_STREAM_END) {
        if (z_.avail_in > 0 || src_->peek() != EOF) inflateReset(&z_);  // next gzip member
        else eof_ = true;
      } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
        throw gzip_error();
      }
      if (got > 0) {
        setg(out_.data(), out_.data(), out_.data() + got);
        return traits_type::to_int_type(*gptr());
      }
      if (eof_ || (z_.avail_in == 0 && !*src_)) return traits_type::eof();
    }
  }

 private:
  std::istream* src_;
  std::vector<char> in_, out_;
  z_stream z_;
  bool ok_, eof_;
};
class filtering_istream : public std::istream {
 public:
  filtering_istream() : std::istream(nullptr), raw_(nullptr), gz_(false) {}
  void push(const gzip_decompressor&) { gz_ = true; }
  void push(std::istream& s) {
    raw_ = &s;
    if (gz_) {
      buf_.attach(s);
      rdbuf(&buf_);
    } else {
      rdbuf(s.rdbuf());
    }
  }

 private:
  std::istream* raw_;
  bool gz_;
  inflate_buf buf_;
};
}  // namespace iostreams
}  // namespace boost
#endif
