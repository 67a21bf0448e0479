// Exhaustive check: sincos_mid == sincos_ref (bitwise) for every float with |y| < 120.
// g++ -O2 -ffp-contract=off -mfma -std=c++17 -pthread tools/sincos_mid_check.cpp -o /tmp/sincos_mid_check && /tmp/sincos_mid_check [stride]
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
#include "../box2d_rs_b200/csrc/b2g_math.h"
int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
  const uint32_t top = 0x42F00000u;  // 120.0f
  const int T = 8;
  std::atomic<long long> bad(0), n(0);
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t]() {
      long long lb = 0, ln = 0;
      for (uint64_t i = (uint64_t)t * stride; i < top; i += (uint64_t)T * stride)
        for (uint32_t sign = 0; sign < 2; ++sign) {
          const uint32_t u = (uint32_t)i | (sign << 31);
          float y; memcpy(&y, &u, 4);
          if (!b2g::sincos_mid_domain(y)) continue;
          float s0, c0, s1, c1;
          b2g::sincos_ref(y, &s0, &c0);
          b2g::sincos_mid(y, &s1, &c1);
          if (memcmp(&s0, &s1, 4) || memcmp(&c0, &c1, 4)) { if (lb < 3) printf("mismatch y=%a ref(%a,%a) mid(%a,%a)\n", y, s0, c0, s1, c1); ++lb; }
          ++ln;
        }
      bad += lb; n += ln;
    });
  for (auto& x : th) x.join();
  printf("checked %lld values, %lld mismatches\n", (long long)n, (long long)bad);
  return bad ? 1 : 0;
}
