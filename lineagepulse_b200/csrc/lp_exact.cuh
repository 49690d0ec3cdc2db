// lp_exact.cuh -- exact, order-independent accumulation of doubles (host + device).
#pragma once
#include "lp_common.cuh"

// ---------------------------------------------------------------------------------------------
// Exact, order-independent summation of the cross-shard quantities.  A double v (|v| < 8e12) is split
// into two integers a = trunc(v 2^20), b = trunc((v 2^20 - a) 2^40), i.e. v = a 2^-20 + b 2^-60 up to
// 2^-60 truncation.  Sums of the a's and b's are plain int64 additions -- associative, so the result
// does not depend on the gene sharding, the chunking or the all-reduce order -- and the total is
// converted back once.  Non-finite inputs are counted in a third word and poison the result.
// ---------------------------------------------------------------------------------------------
#define LP_PI_BLOCK 16   // genes per fixed summation block (shards and chunks start at multiples of this)
struct ExactSum {
    long long a, b, bad;
};
LP_HD void lp_exact_add(ExactSum &s, double v) {
    if (!(fabs(v) < 8.0e12)) {   // NaN, Inf or beyond the int64 range of v * 2^20
        s.bad += 1;
        return;
    }
    const double scaled = v * 1048576.0;            // 2^20, exact
    const double ai = trunc(scaled);
    const double r = scaled - ai;                   // exact, |r| < 1
    s.a += (long long)ai;
    s.b += (long long)trunc(r * 1099511627776.0);   // 2^40
}
LP_HD double lp_exact_value(long long a, long long b, long long bad) {
    if (bad != 0) return NAN;
    return (double)a * 9.5367431640625e-07 + (double)b * 8.6736173798840355e-19;   // 2^-20, 2^-60
}

