// Host-only part of the C ABI: the CPython-exact sampler (lib/ransac/ransac.py:62 uses random.shuffle on the global
// MT19937).  A plain C++ translation unit on purpose: it goes straight to the host compiler (the same loop compiled
// through nvcc's host pass ran 3x slower).
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/sfm_b200.h"

int sfm_internal_set_error(int code, const char* message);  // sfm_api.cu

namespace {

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return sfm_internal_set_error(code, buf);
}

// ---------------------------------------------------------------------------------------
// CPython random: MT19937 + shuffle()  (lib/ransac/ransac.py:62 uses random.shuffle)
// ---------------------------------------------------------------------------------------
struct PyMT {
    uint32_t mt[624];
    uint32_t out[624];  // tempered outputs of the current block (filled per twist: straight-line, vectorisable)
    int pos;
    static inline uint32_t twist(uint32_t u, uint32_t v, uint32_t m) {
        const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
        return m ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    }
    void temper_block() {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = mt[k];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= y >> 18;
            out[k] = y;
        }
    }
    void regenerate() {
        int k = 0;
        for (; k < 624 - 397; ++k) mt[k] = twist(mt[k], mt[k + 1], mt[k + 397]);
        for (; k < 623; ++k) mt[k] = twist(mt[k], mt[k + 1], mt[k - (624 - 397)]);
        mt[623] = twist(mt[623], mt[0], mt[396]);
        temper_block();
        pos = 0;
    }
    void load(const uint32_t* state625) {
        memcpy(mt, state625, 624 * sizeof(uint32_t));
        pos = (int)state625[624];
        temper_block();
    }
    inline uint32_t next() {
        if (pos >= 624) regenerate();
        return out[pos++];
    }
};


// cumulative Fisher-Yates iterations on `perm` (random.shuffle, CPython semantics); row `it` of `table` receives the
// first 8 entries after iteration `it`.  Two phases per iteration: the swap partners j_(n-1), ..., j_1 do not depend on
// the permutation (j = _randbelow(i + 1) = the first getrandbits(bit_length(i + 1)) that is <= i; every stream word is
// consumed, a rejected one just does not advance i), so they are drawn first - loop-carried chain: compare -> subtract,
// words straight from the tempered block, shift constant while i + 1 keeps its bit length - and the swaps follow as a
// plain pass (1.6-2x faster than the fused loop, whose loads and stores alias through the data-dependent index).
static void mt_shuffle_rounds(PyMT& g, int32_t* pp, int64_t n, int64_t h, int32_t* table, int64_t perm_at, int32_t* perm_out,
                              int64_t snap_stride = 0, uint32_t* snap_states = nullptr, int32_t* snap_perms = nullptr) {
    std::vector<uint32_t> jbuf((size_t)n + 8);
    for (int64_t it = 0; it < h; ++it) {
        if (snap_stride > 0 && it % snap_stride == 0) {  // (state, permutation) BEFORE iteration `it`
            const int64_t k = it / snap_stride;
            memcpy(snap_states + 625 * k, g.mt, 624 * sizeof(uint32_t));
            snap_states[625 * k + 624] = (uint32_t)g.pos;
            memcpy(snap_perms + n * k, pp, (size_t)n * sizeof(int32_t));
        }
        uint32_t i = (uint32_t)n - 1;
        uint32_t* jo = jbuf.data();
        while (i >= 1) {
            const int sh = __builtin_clz(i + 1);
            uint32_t lo = (1u << (31 - sh)) - 1u;  // smallest i whose i + 1 has the same bit length
            if (lo < 1) lo = 1;
            while (i >= lo) {
                if (g.pos >= 624) g.regenerate();
                const uint32_t* w = g.out + g.pos;
                const int avail = 624 - g.pos;
                int k = 0;
                for (; k < avail && i >= lo; ++k) {
                    const uint32_t r = w[k] >> sh;
                    const uint32_t acc = r <= i ? 1u : 0u;
                    *jo = r;  // overwritten by the next word unless accepted
                    jo += acc;
                    i -= acc;
                }
                g.pos += k;
            }
        }
        const uint32_t* jp = jbuf.data();
        for (uint32_t q = (uint32_t)n - 1; q >= 1; --q) {
            const uint32_t j = *jp++;
            const int32_t t = pp[q];
            pp[q] = pp[j];
            pp[j] = t;
        }
        if (table) memcpy(table + 8 * it, pp, 8 * sizeof(int32_t));
        if (perm_out && it == perm_at) memcpy(perm_out, pp, (size_t)n * sizeof(int32_t));
    }
}

static int mt_check(const uint32_t* state625, int64_t n, int64_t h) {
    if (!state625) return fail(SFM_ERR_ARG, "null argument");
    if (n < 8 || n > 0x7fffffff) return fail(SFM_ERR_ARG, "need 8 <= n < 2^31 correspondences, got %lld", (long long)n);
    if (h < 0) return fail(SFM_ERR_ARG, "negative hypothesis count");
    if ((int)state625[624] < 0 || (int)state625[624] > 624) return fail(SFM_ERR_ARG, "bad MT position %d", (int)state625[624]);
    return 0;
}

}  // namespace

extern "C" {

int sfm_mt_shuffle_table(uint32_t* state625, int64_t n, int64_t h, int32_t* table, int64_t perm_at,
                         int32_t* perm_out) {
    if (!table) return fail(SFM_ERR_ARG, "null argument");
    if (int r = mt_check(state625, n, h)) return r;
    PyMT g;
    g.load(state625);
    std::vector<int32_t> perm((size_t)n);
    for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = (int32_t)i;
    mt_shuffle_rounds(g, perm.data(), n, h, table, perm_at, perm_out);
    memcpy(state625, g.mt, 624 * sizeof(uint32_t));
    state625[624] = (uint32_t)g.pos;
    return 0;
}

int sfm_mt_shuffle_resume(uint32_t* state625, int64_t n, int64_t h, int32_t* table, int32_t* perm_inout) {
    if (!perm_inout) return fail(SFM_ERR_ARG, "null permutation");
    if (int r = mt_check(state625, n, h)) return r;
    PyMT g;
    g.load(state625);
    mt_shuffle_rounds(g, perm_inout, n, h, table, -1, nullptr);
    memcpy(state625, g.mt, 624 * sizeof(uint32_t));
    state625[624] = (uint32_t)g.pos;
    return 0;
}

int sfm_mt_shuffle_snapshots(uint32_t* state625, int64_t n, int64_t h, int32_t* table, int64_t stride,
                             uint32_t* snap_states, int32_t* snap_perms) {
    if (!table || !snap_states || !snap_perms || stride < 1) return fail(SFM_ERR_ARG, "bad snapshot arguments");
    if (int r = mt_check(state625, n, h)) return r;
    PyMT g;
    g.load(state625);
    std::vector<int32_t> perm((size_t)n);
    for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = (int32_t)i;
    mt_shuffle_rounds(g, perm.data(), n, h, table, -1, nullptr, stride, snap_states, snap_perms);
    memcpy(state625, g.mt, 624 * sizeof(uint32_t));
    state625[624] = (uint32_t)g.pos;
    return 0;
}

}  // extern "C"
