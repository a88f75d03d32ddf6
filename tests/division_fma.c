/* CPU restatement of ddiv_rcp (csrc/gpr_device.cuh): RN(x / y) from z = RN(1 / y) with two FMA corrections, checked against
 * the machine's IEEE division.  Test infrastructure (tests/test_division_fma.py).  Prints the number of mismatches. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t nxt(void) {
    uint64_t a = s[0], b = s[1];
    s[0] = b;
    a ^= a << 23;
    s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}
static inline double u01(void) { return (double)(nxt() >> 11) * (1.0 / 9007199254740992.0); }

static inline double div_rcp(double x, double y, double z) {
    if (!(fabs(x) > 1e-200 && y < 1e150)) return x / y;
    const double q0 = x * z;
    const double r0 = fma(-q0, y, x);
    const double q1 = fma(r0, z, q0);
    const double r1 = fma(-q1, y, x);
    return fma(r1, z, q1);
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 10000000L;
    long bad = 0;
    const double dts[4] = {0.001, 0.002, 0.0005, 1.0 / 3.0};
    for (long i = 0; i < n; ++i) {
        /* the kernel's two uses: (a) x / dt with the host's RN(1/dt); x = difference of O(1..100) quantities, any exponent */
        const double dt = dts[i & 3], zdt = 1.0 / dt;
        const double x = (u01() * 2.0 - 1.0) * ldexp(1.0, (int)(nxt() % 80) - 60);
        if (div_rcp(x, dt, zdt) != x / dt) ++bad;
        /* (b) t / |t| with |t| >= max (velocity 2, acceleration 10 ... jerk limits), |component| <= norm */
        const double y = 0.5 + u01() * ldexp(1.0, (int)(nxt() % 12));
        const double z = 1.0 / y;
        const double c = (u01() * 2.0 - 1.0) * y;
        if (div_rcp(c, y, z) != c / y) ++bad;
        /* (c) operands one ulp apart from representable quotients: x = q * y rounded, the hard cases of division */
        const double q = 1.0 + u01();
        const double xq = q * y;
        if (div_rcp(xq, y, z) != xq / y) ++bad;
        if (div_rcp(nextafter(xq, 4.0 * xq), y, z) != nextafter(xq, 4.0 * xq) / y) ++bad;
    }
    /* zeros keep their sign, tiny and huge operands take the plain division */
    if (signbit(div_rcp(-0.0, 0.001, 1000.0)) == 0) ++bad;
    if (div_rcp(0.0, 0.001, 1000.0) != 0.0) ++bad;
    if (div_rcp(1e-300, 3.0, 1.0 / 3.0) != 1e-300 / 3.0) ++bad;
    if (div_rcp(1.0, 1e200, 1e-200) != 1.0 / 1e200) ++bad;
    printf("%ld\n", bad);
    return 0;
}
