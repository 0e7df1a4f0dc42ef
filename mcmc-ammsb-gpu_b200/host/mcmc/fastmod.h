// mcmc/fastmod.h -- exact `a % d` for a divisor that is fixed over many calls, without a
// hardware divide (Lemire, Kaser, Kurz: "Faster remainder by direct computation", 2019).
//
// The host hot loops are full of remainders by a slowly-changing divisor: the cuckoo bin of a
// key (reference cuckoo.cc:199-209: two 64-bit `%` per set per lookup) and the bucket of a key
// in std::unordered_set (one `%` per insert, one per element per rehash).  A 64-bit `div` costs
// 25-90 cycles and does not pipeline; the multiply form below is four pipelined multiplies.
// The result is the exact remainder for every 64-bit a and d >= 1 (128 fractional bits cover
// 64-bit numerators and 64-bit divisors), so nothing observable changes.
#ifndef MCMC_B200_FASTMOD_H_
#define MCMC_B200_FASTMOD_H_

#include <cstdint>

namespace mcmc {

class FastMod64 {
 public:
  explicit FastMod64(uint64_t d = 1) { Set(d); }
  void Set(uint64_t d) {
    d_ = d;
    // ceil(2^128 / d); wraps to 0 for d == 1, for which Mod() then yields 0 as it should
    m_ = ~static_cast<unsigned __int128>(0) / d + 1;
  }
  uint64_t divisor() const { return d_; }
  uint64_t Mod(uint64_t a) const {
    const unsigned __int128 low = m_ * a;  // fractional part of a / d, 128 bits
    // floor(low * d / 2^128): the top 64 bits of a 128 x 64 -> 192 bit product
    const unsigned __int128 bottom = static_cast<unsigned __int128>(static_cast<uint64_t>(low)) * d_;
    const unsigned __int128 top = (low >> 64) * d_;
    return static_cast<uint64_t>((top + (bottom >> 64)) >> 64);
  }

 private:
  unsigned __int128 m_;
  uint64_t d_;
};

}  // namespace mcmc

#endif  // MCMC_B200_FASTMOD_H_
