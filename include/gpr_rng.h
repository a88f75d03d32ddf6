/*
 * gpr_rng.h — the counter-based random streams of the simulator (shared by the CUDA kernels and the CPU oracle).
 *
 * The reference draws from NumPy generators (PCG64 + ziggurat): `np_random.uniform` for start/goal sampling
 * (planning:377,405; pushing:388-409) and `rng_noise.normal` for sensor noise (basic_envs.py:828,841,855).  A per-env
 * sequential generator cannot be reproduced in a batched kernel, so RNG *parity* with the reference is impossible by
 * construction; what is kept is the *distribution* (uniform over the same box, N(0, sigma^2)).
 *
 * This header is therefore a specification of its own, not a restatement of reference code:
 *   - generator: Philox4x32 (Salmon et al., SC'11), key = 64-bit seed, counter = (env, event, stream, lane); 10 rounds for
 *     the sensor-noise streams, 7 rounds for the rejection-sampling streams (see gpr_rng_block_sampling);
 *   - uniform coordinate: 32 random bits * 2^-32 in [0,1), `low + (high-low)*u` with separate multiply and add
 *     (numpy's `uniform` does the same with 53 bits; 32 bits resolve 1e-10 m and let one block feed two attempts);
 *   - normal float: Box-Muller whose log / sin / cos are fixed polynomials evaluated with IEEE float32 operations only
 *     (explicit fma, correctly-rounded div and sqrt), so the CPU oracle and the GPU produce BIT-IDENTICAL noise and
 *     noisy-mode parity tests can still demand exact collision flags.
 *
 * Every function is `static inline` and compiles as C99, C++ and CUDA.
 */
#ifndef GPR_RNG_H_
#define GPR_RNG_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define GPR_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define GPR_HD static inline
#endif

/* ---- exact float32 primitives (no contraction, no fast-math substitution) ------------------------------------------- */
#if defined(__CUDA_ARCH__)
#define GPR_FMUL(a, b) __fmul_rn((a), (b))
#define GPR_FADD(a, b) __fadd_rn((a), (b))
#define GPR_FSUB(a, b) __fsub_rn((a), (b))
#define GPR_FFMA(a, b, c) __fmaf_rn((a), (b), (c))
#define GPR_FDIV(a, b) __fdiv_rn((a), (b))
#define GPR_FSQRT(a) __fsqrt_rn((a))
#else
/* host: compile with -ffp-contract=off (oracle/Makefile does); fmaf/sqrtf/'/' are correctly rounded on x86-64 SSE */
#define GPR_FMUL(a, b) ((float)(a) * (float)(b))
#define GPR_FADD(a, b) ((float)(a) + (float)(b))
#define GPR_FSUB(a, b) ((float)(a) - (float)(b))
#define GPR_FFMA(a, b, c) fmaf((a), (b), (c))
#define GPR_FDIV(a, b) ((float)(a) / (float)(b))
#define GPR_FSQRT(a) sqrtf((a))
#endif

/* ---- stream ids (counter word 2) ----------------------------------------------------------------------------------- */
/* per control cycle `cyc` and mover lane: word2 = cyc * 4 + block */
#define GPR_RNG_BLOCK_VEL_WALL 0u /* normals: [vel_x, vel_y, wall_x, wall_y]          (planning:430; basic_envs.py:1888) */
#define GPR_RNG_BLOCK_MOVER 1u    /* normals: [mover_x, mover_y, -, -]                (basic_envs.py:1895)               */
#define GPR_RNG_BLOCK_WALL_QUAT 2u  /* box shape: noise on the quaternion used by the wall check  (basic_envs.py:828)    */
#define GPR_RNG_BLOCK_MOVER_QUAT 3u /* box shape: noise on the quaternion used by the mover check                        */
#define GPR_RNG_OBS 0x40000000u        /* normals: [pos_x, pos_y, vel_x, vel_y]       (planning:554-555)                 */
#define GPR_RNG_RESET_CHECK 0x40000001u      /* [wall_x, wall_y, mover_x, mover_y]    (basic_envs.py:1799-1805)          */
#define GPR_RNG_RESET_CHECK_WQUAT 0x40000002u
#define GPR_RNG_RESET_CHECK_MQUAT 0x40000003u
#define GPR_RNG_OBJECT 0x40000004u     /* pushing: normals [obj_x, obj_y, -, -]       (pushing:565)                      */
/* rejection sampling: attempt t of kind k (0 start, 1 goal) reads block (t >> 1) of stream BASE + 2*(t >> 1) + k and
   takes words (0,1) for even t, (2,3) for odd t as (x, y): one Philox block feeds two attempts of one mover */
#define GPR_RNG_RESET_SAMPLE 0x80000000u /* planning:377,405 / pushing:388 */
#define GPR_RNG_RESET_OBJECT 0xC0000000u /* pushing: k = 0 object start, 1 object goal (pushing:401,409) */
/* largest |value| gpr_normal_pair can return: sqrt(-2 ln 2^-24) = 5.7677..., polynomial error < 1e-6 */
#define GPR_NORMAL_ABS_MAX 5.77

typedef struct gpr_u32x4 {
    uint32_t v[4];
} gpr_u32x4;

GPR_HD void gpr_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *lo = (uint32_t)p;
    *hi = (uint32_t)(p >> 32);
#endif
}

/* Philox4x32-R. counter = (c0,c1,c2,c3), key = (k0,k1).  R = 10 is the generator's standard strength; R = 7 is the reduced
 * variant Random123 ships and tests as well (it passes BigCrush; known-answer vectors for both in the test-suite). */
#define GPR_PHILOX_ROUNDS_SAMPLING 7
GPR_HD gpr_u32x4 gpr_philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, const int rounds) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < rounds; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        gpr_mulhilo32(M0, c0, &hi0, &lo0);
        gpr_mulhilo32(M1, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    gpr_u32x4 out;
    out.v[0] = c0;
    out.v[1] = c1;
    out.v[2] = c2;
    out.v[3] = c3;
    return out;
}

GPR_HD gpr_u32x4 gpr_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return gpr_philox4x32_r(c0, c1, c2, c3, k0, k1, 10);
}

/* One random block for (env, event, stream, lane) under a 64-bit seed: the sensor-noise streams. */
GPR_HD gpr_u32x4 gpr_rng_block(uint64_t seed, uint32_t env_global, uint32_t event, uint32_t stream, uint32_t lane) {
    return gpr_philox4x32_10(env_global, event, stream, lane, (uint32_t)seed, (uint32_t)(seed >> 32));
}
/* The same for the REJECTION-SAMPLING streams (GPR_RNG_RESET_SAMPLE / GPR_RNG_RESET_OBJECT): Philox4x32-7.  Acceptance of
 * the reference's all-movers-at-once sampling is ~1 % (planning:369-385), so a reset consumes hundreds of blocks and the
 * generator is a third of the auto-reset kernel's instructions; the 7-round variant still passes BigCrush (Salmon et al.,
 * SC'11, table 2) and saves 30 % of them.  (Different streams of one counter space never overlap: the stream id is a
 * counter word.) */
GPR_HD gpr_u32x4 gpr_rng_block_sampling(uint64_t seed, uint32_t env_global, uint32_t event, uint32_t stream, uint32_t lane) {
    return gpr_philox4x32_r(env_global, event, stream, lane, (uint32_t)seed, (uint32_t)(seed >> 32), GPR_PHILOX_ROUNDS_SAMPLING);
}

/* 53-bit uniform in [0,1) from two words. */
GPR_HD double gpr_uniform53(uint32_t hi, uint32_t lo) {
    uint64_t bits = (((uint64_t)hi << 32) | (uint64_t)lo) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

/* 32-bit uniform in [0,1): the resolution of sampled start / goal coordinates (span * 2^-32 ~ 1e-10 m). */
GPR_HD double gpr_uniform32(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

/* coordinates of rejection-sampling attempt t for (env, event, kind, lane): low + span * u, multiply and add rounded
   separately like numpy's Generator.uniform */
GPR_HD void gpr_sample_xy(uint64_t seed, uint32_t env_global, uint32_t event, uint32_t base, uint32_t kind, uint32_t t,
                          uint32_t lane, double* ux, double* uy) {
    gpr_u32x4 r = gpr_rng_block_sampling(seed, env_global, event, base + 2u * (t >> 1) + kind, lane);
    *ux = gpr_uniform32(r.v[2u * (t & 1u)]);
    *uy = gpr_uniform32(r.v[2u * (t & 1u) + 1u]);
}

/* Two independent standard normals from two words (Box-Muller, fixed float32 polynomials). */
GPR_HD void gpr_normal_pair(uint32_t ra, uint32_t rb, float* n0, float* n1) {
    /* u1 in (0,1], 24 bits; exactly representable */
    float u1 = GPR_FMUL((float)((ra >> 8) + 1u), 5.9604644775390625e-8f);
    /* ---- ln(u1): u1 = m * 2^e, m in [sqrt(1/2), sqrt(2)) ---- */
    union {
        float f;
        uint32_t u;
    } cvt;
    cvt.f = u1;
    int32_t e = (int32_t)(cvt.u >> 23) - 127;
    cvt.u = (cvt.u & 0x007FFFFFu) | 0x3F800000u; /* m in [1,2) */
    float m = cvt.f;
    if (m > 1.41421356f) {
        m = GPR_FMUL(m, 0.5f);
        e += 1;
    }
    float z = GPR_FDIV(GPR_FSUB(m, 1.0f), GPR_FADD(m, 1.0f)); /* |z| <= 0.1716 */
    float z2 = GPR_FMUL(z, z);
    /* ln(m) = 2 z (1 + z^2/3 + z^4/5 + z^6/7 + z^8/9) */
    float p = 0.1111111111f;
    p = GPR_FFMA(p, z2, 0.1428571429f);
    p = GPR_FFMA(p, z2, 0.2f);
    p = GPR_FFMA(p, z2, 0.3333333333f);
    p = GPR_FFMA(p, z2, 1.0f);
    float lnm = GPR_FMUL(GPR_FADD(z, z), p);
    float lnu = GPR_FFMA((float)e, 0.6931471806f, lnm); /* <= 0 */
    float radius = GPR_FSQRT(GPR_FMUL(-2.0f, lnu));
    /* ---- direction: quadrant from the top 2 bits, theta in [-pi/4, pi/4) from the next 22 ---- */
    uint32_t quad = rb >> 30;
    float f = GPR_FSUB(GPR_FMUL((float)((rb >> 8) & 0x003FFFFFu), 2.384185791015625e-7f), 0.5f); /* [-0.5, 0.5) */
    float th = GPR_FMUL(f, 1.5707963268f);
    float t2 = GPR_FMUL(th, th);
    /* sin: th (1 - t2/6 + t2^2/120 - t2^3/5040 + t2^4/362880) */
    float sp = 2.7557319224e-6f;
    sp = GPR_FFMA(sp, t2, -1.9841269841e-4f);
    sp = GPR_FFMA(sp, t2, 8.3333333333e-3f);
    sp = GPR_FFMA(sp, t2, -0.1666666667f);
    sp = GPR_FFMA(sp, t2, 1.0f);
    float s = GPR_FMUL(th, sp);
    /* cos: 1 - t2/2 + t2^2/24 - t2^3/720 + t2^4/40320 - t2^5/3628800 */
    float cp = -2.7557319224e-7f;
    cp = GPR_FFMA(cp, t2, 2.4801587302e-5f);
    cp = GPR_FFMA(cp, t2, -1.3888888889e-3f);
    cp = GPR_FFMA(cp, t2, 4.1666666667e-2f);
    cp = GPR_FFMA(cp, t2, -0.5f);
    float c = GPR_FFMA(cp, t2, 1.0f);
    /* rotate by quad * 90 degrees */
    float cx = (quad == 0u) ? c : (quad == 1u) ? -s : (quad == 2u) ? -c : s;
    float sx = (quad == 0u) ? s : (quad == 1u) ? c : (quad == 2u) ? -s : -c;
    *n0 = GPR_FMUL(radius, cx);
    *n1 = GPR_FMUL(radius, sx);
}

/* Four standard normals for (env, event, stream, lane). */
GPR_HD void gpr_normal4(uint64_t seed, uint32_t env_global, uint32_t event, uint32_t stream, uint32_t lane,
                        float out[4]) {
    gpr_u32x4 r = gpr_rng_block(seed, env_global, event, stream, lane);
    gpr_normal_pair(r.v[0], r.v[1], &out[0], &out[1]);
    gpr_normal_pair(r.v[2], r.v[3], &out[2], &out[3]);
}

#endif /* GPR_RNG_H_ */
