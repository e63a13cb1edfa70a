// Response text builder: the subset of std::ostream's operator<< the handlers use, with the same output bytes
// (integers in decimal, float / double as "%g" with precision 6 -- what num_put produces for a default-constructed
// stream) and none of the locale machinery.  std::to_chars(general, 6) is specified to equal printf("%.6g") in the
// "C" locale.
#ifndef CKM_HOST_TEXT_H
#define CKM_HOST_TEXT_H
#include <charconv>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

namespace ckm_text {

class Text {
public:
    Text() { buf_.reserve(1 << 16); }
    Text &operator<<(const char *s) {
        buf_.append(s);
        return *this;
    }
    Text &operator<<(const std::string &s) {
        buf_.append(s);
        return *this;
    }
    Text &operator<<(char c) {
        buf_.push_back(c);
        return *this;
    }
    Text &operator<<(int v) { return put_int((long long)v); }
    Text &operator<<(long v) { return put_int((long long)v); }
    Text &operator<<(long long v) { return put_int(v); }
    Text &operator<<(unsigned v) { return put_uint((unsigned long long)v); }
    Text &operator<<(unsigned long v) { return put_uint((unsigned long long)v); }
    Text &operator<<(unsigned long long v) { return put_uint(v); }
    Text &operator<<(unsigned short v) { return put_uint((unsigned long long)v); }
    Text &operator<<(short v) { return put_int((long long)v); }
    Text &operator<<(float v) { return put_double((double)v); }  // num_put formats a float through double
    Text &operator<<(double v) { return put_double(v); }
    const std::string &str() const { return buf_; }
    // malloc'ed NUL-terminated copy (ckm_free_text releases it)
    char *dup() const {
        char *p = (char *)malloc(buf_.size() + 1);
        if (!p) return nullptr;
        memcpy(p, buf_.data(), buf_.size() + 1);
        return p;
    }

private:
    Text &put_int(long long v) {
        char tmp[24];
        auto r = std::to_chars(tmp, tmp + sizeof tmp, v);
        buf_.append(tmp, r.ptr);
        return *this;
    }
    Text &put_uint(unsigned long long v) {
        char tmp[24];
        auto r = std::to_chars(tmp, tmp + sizeof tmp, v);
        buf_.append(tmp, r.ptr);
        return *this;
    }
    Text &put_double(double v) {
        char tmp[64];
        auto r = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::general, 6);
        buf_.append(tmp, r.ptr);
        return *this;
    }
    std::string buf_;
};

}  // namespace ckm_text
#endif
