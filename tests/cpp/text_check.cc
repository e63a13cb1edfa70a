// ckm_text::Text prints numbers exactly like a default std::ostream (what the reference's handlers use).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#include <sstream>

#include "../../close_kmers_b200/host/text.h"

template <class T>
static bool same(T v) {
    std::ostringstream os;
    os << v;
    ckm_text::Text t;
    t << v;
    if (os.str() != t.str()) {
        fprintf(stderr, "mismatch: ostream '%s' text '%s'\n", os.str().c_str(), t.str().c_str());
        return false;
    }
    return true;
}

int main() {
    std::mt19937_64 rng(12345);
    unsigned long bad = 0, n = 0;
    const float specials[] = {0.0f, -0.0f, 1.0f, -1.0f, 0.1f, 0.5f, 1e-5f, 9.99999e-5f, 1e-4f, 123456.0f, 999999.5f, 1e6f, 1234567.0f, 1e-38f,
                              3.4e38f, 1.17549435e-38f, 1e-45f, 0.333333343f, 2.0f / 3.0f, 100000.0f, 99999.95f, 0.000123456789f,
                              std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(), 54.6202f, 0.783333f};
    for (float f : specials) bad += !same(f), n++;
    for (int i = 0; i < 3000000; i++) {  // every exponent, random mantissas
        uint32_t bits = (uint32_t)rng();
        float f;
        memcpy(&f, &bits, 4);
        if (std::isnan(f)) continue;
        bad += !same(f), n++;
    }
    for (int i = 0; i < 1000000; i++) {  // the ranges scores live in: sums of reciprocals, small ratios
        float f = (float)(rng() % 100000) / (float)(1 + rng() % 997);
        bad += !same(f), n++;
        bad += !same(f * 1e-4f), n++;
    }
    for (int i = 0; i < 200000; i++) {
        bad += !same((int)(int32_t)rng()), n++;
        bad += !same((unsigned)rng()), n++;
        bad += !same((unsigned long)rng()), n++;
        bad += !same((long)rng()), n++;
        bad += !same((unsigned short)rng()), n++;
        bad += !same((double)(rng() % 1000000) / 7.0), n++;
    }
    printf("%lu values, %lu mismatches\n", n, bad);
    return bad != 0;
}
