// kernels.cu -- the hot path of filter_variants_pipeline as sm_100a CUDA kernels.
//
//   K0  line index     : newline scan of the VCF text -> line_start[], n_records
//   K1  field parse    : one thread per record walks the 8 fixed columns, the INFO
//                        key=value list and the FORMAT/sample pair, decodes the tags
//                        the plan needs with htslib's typing rules (int32 / float32 /
//                        dictionary-encoded strings) into a shared-memory slot tile that
//                        is written out as coalesced columnar rows raw[slot][record]
//                        (replaces vcftools.py:63-89 + :196-214 of the reference)
//   K2  feature assembly: raw slots -> fp32 feature matrix feats[feature][record] with the
//                        fitted transformer's missing/absent policies (transformers.py:221-344)
//   K3  inference      : feature tile staged in shared memory, tree-ensemble / logistic
//                        evaluation in the reference library's own arithmetic order,
//                        fp64 phred/qual math and the FILTER decision fused in
//                        (variant_filtering_utils.py:123-124, filter_variants_pipeline.py:170-195)
//
// No tensor cores: there is no dense contraction on this path; the kernels are
// byte/integer work bounded by HBM and by instruction issue.
#include "kernels.cuh"

#include <math.h>

#define WARP 32
#define K1_TPB 128
#define K2_TPB 256
#define K3_TPB 128

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned nl_count4(unsigned w) { return __popc(__vcmpeq4(w, 0x0A0A0A0Au)) >> 3; }
__device__ __forceinline__ unsigned nl_bits4(unsigned w) {
    unsigned m = __vcmpeq4(w, 0x0A0A0A0Au);
    return ((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u);
}
__device__ __forceinline__ uint4 ld_stream16(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------
// K0: line index
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k0_count(const uint8_t* __restrict__ text, size_t n_bytes,
                                                uint32_t* __restrict__ chunk_count, size_t n_chunks) {
    const int lane = threadIdx.x & 31;
    size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t c = warp; c < n_chunks; c += n_warps) {
        const size_t base = c * K0_CHUNK_BYTES;
        unsigned cnt = 0;
#pragma unroll
        for (int it = 0; it < K0_CHUNK_BYTES / 512; ++it) {
            const size_t off = base + (size_t)it * 512 + (size_t)lane * 16;
            if (off + 16 <= n_bytes) {
                uint4 v = ld_stream16(reinterpret_cast<const uint4*>(text + off));
                cnt += nl_count4(v.x) + nl_count4(v.y) + nl_count4(v.z) + nl_count4(v.w);
            } else {
                for (size_t q = off; q < n_bytes && q < off + 16; ++q) cnt += (text[q] == '\n');
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        if (lane == 0) chunk_count[c] = cnt;
    }
}

// single block: exclusive scan of chunk counts in place; publishes n_records
__global__ void __launch_bounds__(1024) k0_scan(uint32_t* __restrict__ chunk, size_t n_chunks,
                                                const uint8_t* __restrict__ text, size_t n_bytes,
                                                int64_t* __restrict__ line_start, size_t cap_records,
                                                int64_t* __restrict__ n_records, unsigned long long* err) {
    __shared__ unsigned long long part[1024];
    const int t = threadIdx.x;
    const size_t per = (n_chunks + 1023) / 1024;
    const size_t lo = (size_t)t * per, hi = min(lo + per, n_chunks);
    unsigned long long s = 0;
    for (size_t i = lo; i < hi; ++i) s += chunk[i];
    part[t] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned long long v = (t >= d) ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = (t == 0) ? 0 : part[t - 1];
    for (size_t i = lo; i < hi; ++i) {
        unsigned c = chunk[i];
        chunk[i] = (uint32_t)run;
        run += c;
    }
    if (t == 1023) {
        long long total = (long long)part[1023];
        if (n_bytes > 0 && text[n_bytes - 1] != '\n') {  // contract: the batch ends with '\n'
            atomicMin(err, ugvc_pack_error(total, 0xFFFF, REASON_MALFORMED_LINE));
        }
        if ((size_t)total > cap_records) {
            atomicMin(err, ugvc_pack_error((long long)cap_records, 0xFFFF, REASON_TOO_MANY_ELEMS));
            total = (long long)cap_records;
        }
        *n_records = total;
        line_start[0] = 0;
    }
}

__global__ void __launch_bounds__(256) k0_write(const uint8_t* __restrict__ text, size_t n_bytes,
                                                const uint32_t* __restrict__ chunk_first, size_t n_chunks,
                                                int64_t* __restrict__ line_start, size_t cap_records) {
    const int lane = threadIdx.x & 31;
    size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t c = warp; c < n_chunks; c += n_warps) {
        const size_t base = c * K0_CHUNK_BYTES;
        size_t rec = chunk_first[c];
#pragma unroll 1
        for (int it = 0; it < K0_CHUNK_BYTES / 512; ++it) {
            const size_t off = base + (size_t)it * 512 + (size_t)lane * 16;
            unsigned mask = 0;
            if (off + 16 <= n_bytes) {
                uint4 v = ld_stream16(reinterpret_cast<const uint4*>(text + off));
                mask = nl_bits4(v.x) | (nl_bits4(v.y) << 4) | (nl_bits4(v.z) << 8) | (nl_bits4(v.w) << 12);
            } else {
                for (size_t q = off; q < n_bytes && q < off + 16; ++q)
                    if (text[q] == '\n') mask |= 1u << (q - off);
            }
            unsigned cnt = __popc(mask), incl = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                unsigned v = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += v;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            size_t idx = rec + (incl - cnt);
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                if (idx + 1 <= cap_records) line_start[idx + 1] = (int64_t)(off + b + 1);
                ++idx;
            }
            rec += total;
        }
    }
}

void launch_k0(const uint8_t* d_text, size_t n_bytes, uint32_t* chunk_first, int64_t* line_start,
               size_t cap_records, int64_t* d_n_records, unsigned long long* d_err, int sm_count,
               cudaStream_t st) {
    const size_t n_chunks = (n_bytes + K0_CHUNK_BYTES - 1) / K0_CHUNK_BYTES;
    const size_t warps_per_block = 8;
    size_t blocks = (n_chunks + warps_per_block - 1) / warps_per_block;
    const size_t max_blocks = (size_t)sm_count * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks == 0) blocks = 1;
    k0_count<<<(unsigned)blocks, 256, 0, st>>>(d_text, n_bytes, chunk_first, n_chunks);
    k0_scan<<<1, 1024, 0, st>>>(chunk_first, n_chunks, d_text, n_bytes, line_start, cap_records, d_n_records, d_err);
    k0_write<<<(unsigned)blocks, 256, 0, st>>>(d_text, n_bytes, chunk_first, n_chunks, line_start, cap_records);
}

// ------------------------------------------------------------------------------------------
// K1: field parse
// ------------------------------------------------------------------------------------------
__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

struct K1Shared {
    const PlanTag* tags;
    const PlanSlot* slots;
    const PlanDict* dicts;
    const PlanString* strings;
    const uint8_t* htab;
    uint32_t* tile;  // [n_slots][K1_TPB]
};

enum { NUM_OK = 0, NUM_MISSING = 1, NUM_BAD = 2 };

__device__ __forceinline__ bool is_digit(unsigned c) { return (c - '0') <= 9u; }
__device__ __forceinline__ unsigned lower(unsigned c) { return c | 0x20u; }

// Parse one numeric token starting at p (htslib: strtod -> float32 for Float, strtol for
// Integer).  On return p is one past the token.  The double is exactly strtod's result for
// every literal with <= 19 significant digits whose decimal exponent fits Clinger's exact
// window; anything else reports NUM_BAD (never a silently different value).
__device__ int parse_num(const uint8_t*& p, double& out) {
    unsigned c = *p;
    bool neg = false;
    if (c == '-' || c == '+') {
        neg = (c == '-');
        c = *++p;
    }
    unsigned long long m = 0;
    int exp10 = 0;
    bool any = false, inexact = false;
    while (is_digit(c)) {
        any = true;
        const unsigned d = c - '0';
        if (m < 1844674407370955161ull) m = m * 10 + d;
        else { ++exp10; inexact |= (d != 0); }
        c = *++p;
    }
    if (c == '.') {
        c = *++p;
        while (is_digit(c)) {
            any = true;
            const unsigned d = c - '0';
            if (m < 1844674407370955161ull) { m = m * 10 + d; --exp10; }
            else inexact |= (d != 0);
            c = *++p;
        }
        if (!any) {  // a lone "." (possibly signed): missing
            out = 0.0;
            return neg ? NUM_BAD : NUM_MISSING;
        }
    }
    if (!any) {
        // nan / inf / infinity (any case), as strtod accepts them
        const unsigned a = lower(c);
        if (a == 'n' && lower(p[1]) == 'a' && lower(p[2]) == 'n') {
            p += 3;
            out = __longlong_as_double(0x7FF8000000000000ll);
            return NUM_OK;
        }
        if (a == 'i' && lower(p[1]) == 'n' && lower(p[2]) == 'f') {
            p += 3;
            if (lower(p[0]) == 'i' && lower(p[1]) == 'n' && lower(p[2]) == 'i' && lower(p[3]) == 't' &&
                lower(p[4]) == 'y')
                p += 5;
            out = neg ? -INFINITY : INFINITY;
            return NUM_OK;
        }
        out = 0.0;
        return NUM_BAD;
    }
    if (lower(c) == 'e') {
        const uint8_t* q = p + 1;
        unsigned e = *q;
        bool eneg = false;
        if (e == '-' || e == '+') {
            eneg = (e == '-');
            e = *++q;
        }
        if (is_digit(e)) {
            int ev = 0;
            while (is_digit(e)) {
                if (ev < 100000) ev = ev * 10 + (int)(e - '0');
                e = *++q;
            }
            exp10 += eneg ? -ev : ev;
            p = q;
        }
    }
    double v;
    if (m == 0) {
        v = 0.0;
    } else if (inexact || m > (1ull << 53)) {
        out = 0.0;
        return NUM_BAD;
    } else if (exp10 == 0) {
        v = (double)m;
    } else if (exp10 > 0 && exp10 <= 22) {
        v = (double)m * c_pow10[exp10];
    } else if (exp10 < 0 && exp10 >= -22) {
        v = (double)m / c_pow10[-exp10];
    } else if (exp10 > 22 && exp10 <= 22 + 15) {
        // m * 10^(exp10-22) may still be an exact integer <= 2^53
        const double scaled = (double)m * c_pow10[exp10 - 22];
        if (scaled > 9007199254740992.0) {
            out = 0.0;
            return NUM_BAD;
        }
        v = scaled * 1e22;
    } else {
        out = 0.0;
        return NUM_BAD;
    }
    out = neg ? -v : v;
    return NUM_OK;
}

__device__ __forceinline__ unsigned base_code(unsigned c) {
    // {A:1, T:2, G:3, C:4}  (transformers.py:72-77)
    return c == 'A' ? 1u : c == 'T' ? 2u : c == 'G' ? 3u : c == 'C' ? 4u : 0u;
}
__device__ __forceinline__ unsigned motif_code(unsigned c) { return c == 'N' ? 5u : base_code(c); }

__device__ __forceinline__ bool bytes_equal(const uint8_t* a, const char* b, int n) {
    for (int i = 0; i < n; ++i)
        if (a[i] != (uint8_t)b[i]) return false;
    return true;
}

__device__ __forceinline__ int find_tag(const K1Shared& sp, const uint8_t* key, int len, unsigned hash) {
    if (len <= 0 || len > UGVC_NAME_MAX) return -1;
    unsigned idx = (hash ^ (hash >> 8) ^ (hash >> 16)) & 255u;
    for (int probe = 0; probe < 256; ++probe) {
        const unsigned t = sp.htab[idx];
        if (t == 0xFFu) return -1;
        const PlanTag& tg = sp.tags[t];
        if (tg.len == len && bytes_equal(key, tg.name, len)) return (int)t;
        idx = (idx + 1) & 255u;
    }
    return -1;
}

__device__ __forceinline__ void store_slot(const K1Shared& sp, int slot, uint32_t bits) {
    sp.tile[slot * K1_TPB + threadIdx.x] = bits;
}
__device__ __forceinline__ void store_slot_f(const K1Shared& sp, int slot, float v) {
    store_slot(sp, slot, __float_as_uint(v));
}

__device__ uint32_t reduce_string(const K1Shared& sp, const PlanSlot& sl, const uint8_t* a, const uint8_t* b) {
    const int n = (int)(b - a);
    switch (sl.reducer) {
        case RED_BASE:
            return __float_as_uint(n == 1 ? (float)base_code(a[0]) : 0.0f);
        case RED_INSDEL:
            if (n == 3 && a[0] == 'i' && a[1] == 'n' && a[2] == 's') return __float_as_uint(-1.0f);
            if (n == 3 && a[0] == 'd' && a[1] == 'e' && a[2] == 'l') return __float_as_uint(1.0f);
            if (n == 2 && a[0] == 'N' && a[1] == 'A') return __float_as_uint(0.0f);
            return RAW_ERR;
        case RED_DICT: {
            const PlanDict d = sp.dicts[sl.dict];
            for (int i = 0; i < d.n_strings; ++i) {
                const PlanString& s = sp.strings[d.first_string + i];
                if (s.len == n && bytes_equal(a, s.s, n)) return __float_as_uint((float)i);
            }
            return RAW_ERR;
        }
        default:
            return RAW_ERR;
    }
}

// Decode one tag value that starts at p; vend is the character that ends a value in this
// column (';' in INFO, ':' in the sample column).  Leaves p on the terminating character.
__device__ void parse_value(const K1Shared& sp, const PlanTag& tg, unsigned kind, const uint8_t*& p,
                            unsigned vend) {
    const int s0 = tg.first_slot, s1 = tg.first_slot + tg.n_slots;
    for (int s = s0; s < s1; ++s) store_slot(sp, s, RAW_MISSING);
    const unsigned type = kind & KIND_TYPE_MASK;
    const bool scalar = (kind & KIND_SCALAR) != 0;
    const uint8_t* vstart = p;
    int e = 0;
    if (type == KIND_FLAG) {
        while (*p != vend && *p != '\t' && *p != '\n') ++p;
        return;
    }
    for (;;) {
        const uint8_t* a = p;
        double val = 0.0;
        int st = NUM_OK;
        if (type == KIND_INT || type == KIND_FLOAT) {
            st = parse_num(p, val);
            unsigned c = *p;
            if (c != ',' && c != vend && c != '\t' && c != '\n') {  // trailing garbage in the token
                st = NUM_BAD;
                while (c != ',' && c != vend && c != '\t' && c != '\n') c = *++p;
            }
        } else {
            unsigned c = *p;
            if (scalar)
                while (c != vend && c != '\t' && c != '\n') c = *++p;
            else
                while (c != ',' && c != vend && c != '\t' && c != '\n') c = *++p;
        }
        for (int s = s0; s < s1; ++s) {
            const PlanSlot sl = sp.slots[s];
            if (sl.elem != e) continue;
            if (sl.reducer == RED_NUM) {
                if (type == KIND_INT) {
                    if (st == NUM_OK) {
                        // htslib keeps int32; features are fp32 downstream
                        store_slot_f(sp, s, (float)(long long)val);
                    } else
                        store_slot(sp, s, st == NUM_MISSING ? RAW_MISSING : RAW_ERR);
                } else if (type == KIND_FLOAT) {
                    if (st == NUM_OK) {
                        const float f = (float)val;  // float32(strtod(text)), round-to-nearest-even
                        store_slot(sp, s, isnan(f) ? RAW_MISSING : __float_as_uint(f));
                    } else
                        store_slot(sp, s, st == NUM_MISSING ? RAW_MISSING : RAW_ERR);
                } else
                    store_slot(sp, s, RAW_ERR);
            } else if (sl.reducer <= RED_DICT) {
                store_slot(sp, s, type == KIND_STR ? reduce_string(sp, sl, a, p) : RAW_ERR);
            }
        }
        ++e;
        if (*p == ',' && !(scalar && type == KIND_STR)) {
            ++p;
            continue;
        }
        break;
    }
    // whole-value reducers
    for (int s = s0; s < s1; ++s) {
        const PlanSlot sl = sp.slots[s];
        if (sl.elem != ELEM_WHOLE) continue;
        switch (sl.reducer) {
            case RED_LEN:
                store_slot_f(sp, s, (float)e);
                break;
            case RED_MOTIF_L:
            case RED_MOTIF_R: {
                // list(x): characters of a str, elements of a tuple (only single-character
                // elements can match a base) -- transformers.py:36-60
                double num = 0.0, scale = 1.0;
                const bool left = sl.reducer == RED_MOTIF_L;
                const uint8_t* q = vstart;
                if (type != KIND_STR) {
                    store_slot(sp, s, RAW_ERR);
                    break;
                }
                while (q < p) {
                    unsigned code;
                    if (scalar) {
                        code = motif_code(*q++);
                    } else {
                        const uint8_t* ea = q;
                        while (q < p && *q != ',') ++q;
                        code = (q - ea == 1) ? motif_code(*ea) : 0u;
                        if (q < p) ++q;  // skip ','
                    }
                    if (left) {
                        num += scale * (double)code;
                        scale *= 10.0;
                    } else
                        num = num * 10.0 + (double)code;
                }
                store_slot_f(sp, s, (float)num);
                break;
            }
            case RED_STRNUM: {
                const uint8_t* q = vstart;
                double v;
                const int st = parse_num(q, v);
                store_slot(sp, s, (st == NUM_OK && q == p && !isnan(v)) ? __float_as_uint((float)v) : RAW_ERR);
                break;
            }
            case RED_GT_HOM: {
                const bool hom = (p - vstart == 3) && vstart[0] == '1' && (vstart[1] == '/' || vstart[1] == '|') &&
                                 vstart[2] == '1';
                store_slot_f(sp, s, hom ? 1.0f : 0.0f);
                break;
            }
            default:
                break;
        }
    }
}

__device__ __forceinline__ const uint8_t* skip_to(const uint8_t* p, unsigned d0) {
    unsigned c = *p;
    while (c != d0 && c != '\t' && c != '\n') c = *++p;
    return p;
}

__global__ void __launch_bounds__(K1_TPB) k1_parse(const __grid_constant__ DevPlan plan,
                                                   const uint8_t* __restrict__ text,
                                                   const int64_t* __restrict__ line_start,
                                                   const int64_t* __restrict__ n_records_p,
                                                   uint32_t* __restrict__ raw, size_t row_stride,
                                                   ugvc_recinfo* __restrict__ recinfo, unsigned long long* err,
                                                   long long* counts) {
    extern __shared__ __align__(16) uint8_t smem[];
    // shared-memory copies of the small plan tables
    const int n_tags = plan.h.n_tags, n_slots = plan.h.n_slots;
    PlanTag* s_tags = reinterpret_cast<PlanTag*>(smem);
    PlanSlot* s_slots = reinterpret_cast<PlanSlot*>(s_tags + n_tags);
    PlanDict* s_dicts = reinterpret_cast<PlanDict*>(s_slots + ((n_slots + 1) & ~1));
    PlanString* s_strings = reinterpret_cast<PlanString*>(s_dicts + ((plan.h.n_dicts + 1) & ~1));
    uint8_t* s_htab = reinterpret_cast<uint8_t*>(s_strings + plan.h.n_dict_strings);
    uint32_t* s_tile = reinterpret_cast<uint32_t*>(s_htab + 256);
    {
        const uint32_t* src;
        uint32_t* dst;
        src = reinterpret_cast<const uint32_t*>(plan.tags);
        dst = reinterpret_cast<uint32_t*>(s_tags);
        for (int i = threadIdx.x; i < n_tags * 8; i += K1_TPB) dst[i] = src[i];
        src = reinterpret_cast<const uint32_t*>(plan.slots);
        dst = reinterpret_cast<uint32_t*>(s_slots);
        for (int i = threadIdx.x; i < n_slots; i += K1_TPB) dst[i] = src[i];
        src = reinterpret_cast<const uint32_t*>(plan.dicts);
        dst = reinterpret_cast<uint32_t*>(s_dicts);
        for (int i = threadIdx.x; i < (int)plan.h.n_dicts; i += K1_TPB) dst[i] = src[i];
        src = reinterpret_cast<const uint32_t*>(plan.strings);
        dst = reinterpret_cast<uint32_t*>(s_strings);
        for (int i = threadIdx.x; i < (int)plan.h.n_dict_strings * 8; i += K1_TPB) dst[i] = src[i];
        for (int i = threadIdx.x; i < 256; i += K1_TPB) s_htab[i] = plan.htab[i];
    }
    K1Shared sp;
    sp.tags = s_tags;
    sp.slots = s_slots;
    sp.dicts = s_dicts;
    sp.strings = s_strings;
    sp.htab = s_htab;
    sp.tile = s_tile;
    __syncthreads();

    const long long n_rec = *n_records_p;
    const long long n_tiles = (n_rec + K1_TPB - 1) / K1_TPB;
    unsigned cg_local = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // every slot starts ABSENT (the reference's defaultdict(lambda: None))
        for (int i = threadIdx.x; i < n_slots * K1_TPB; i += K1_TPB) s_tile[i] = RAW_ABSENT;
        __syncthreads();
        const long long rec = tile * K1_TPB + threadIdx.x;
        if (rec < n_rec) {
            const int64_t ls = line_start[rec];
            const uint8_t* const line = text + ls;
            const uint8_t* p = line;
            ugvc_recinfo ri;
            ri.flags = 0;
            bool malformed = false;
            // ---- CHROM
            p = skip_to(p, '\t');
            malformed |= (*p != '\t');
            // ---- POS
            long long pos = 0;
            if (!malformed) {
                ++p;
                unsigned c = *p;
                while (is_digit(c)) {
                    pos = pos * 10 + (c - '0');
                    c = *++p;
                }
                malformed |= (*p != '\t');
            }
            ri.pos = (int32_t)pos;
            // ---- ID
            if (!malformed) {
                p = skip_to(p + 1, '\t');
                malformed |= (*p != '\t');
            }
            // ---- REF / ALT -> allele codes, indel flag, CG flag
            float a0 = 0.f, a1 = 0.f;
            bool a1_missing = true, indel = false, cg = false;
            int n_alleles = 0;
            if (!malformed) {
                const uint8_t* ra = p + 1;
                p = skip_to(ra, '\t');
                const int ref_len = (int)(p - ra);
                a0 = ref_len == 1 ? (float)base_code(ra[0]) : 0.f;
                cg |= ref_len == 3 && ((ra[0] == 'G' && ra[1] == 'G' && ra[2] == 'C') ||
                                       (ra[0] == 'C' && ra[1] == 'C' && ra[2] == 'G'));
                n_alleles = 1;
                ri.flags |= (unsigned)(ref_len > 0xFFFFFF ? 0xFFFFFF : ref_len) << 8;
                malformed |= (*p != '\t');
                if (!malformed) {
                    const uint8_t* aa = p + 1;
                    if (aa[0] == '.' && (aa[1] == '\t' || aa[1] == '\n')) {
                        p = aa + 1;  // ALT "." -> alleles == (REF,)
                    } else {
                        for (;;) {
                            p = skip_to(aa, ',');
                            const int alen = (int)(p - aa);
                            if (n_alleles == 1) {
                                a1 = alen == 1 ? (float)base_code(aa[0]) : 0.f;
                                a1_missing = false;
                            }
                            indel |= (alen != ref_len);
                            cg |= alen == 3 && ((aa[0] == 'G' && aa[1] == 'G' && aa[2] == 'C') ||
                                                (aa[0] == 'C' && aa[1] == 'C' && aa[2] == 'G'));
                            ++n_alleles;
                            if (*p != ',') break;
                            aa = p + 1;
                        }
                    }
                    malformed |= (*p != '\t');
                }
            }
            // ---- QUAL
            uint32_t qual_bits = RAW_MISSING;
            unsigned off;
            off = (unsigned)(p + 1 - line);
            ri.qual_off = off > 0xFFFFu ? 0xFFFFu : (uint16_t)off;
            if (!malformed) {
                ++p;
                if (p[0] == '.' && p[1] == '\t') {
                    ++p;
                } else {
                    double qv;
                    const int st = parse_num(p, qv);
                    if (st == NUM_OK && *p == '\t') {
                        const float f = (float)qv;
                        qual_bits = isnan(f) ? RAW_MISSING : __float_as_uint(f);
                    } else {
                        qual_bits = RAW_ERR;
                        p = skip_to(p, '\t');
                    }
                }
                malformed |= (*p != '\t');
            }
            // ---- FILTER (kept as bytes; only its position is reported)
            off = (unsigned)(p + 1 - line);
            ri.filter_off = off > 0xFFFFu ? 0xFFFFu : (uint16_t)off;
            if (!malformed) {
                p = skip_to(p + 1, '\t');
                malformed |= (*p != '\t');
            }
            // ---- INFO
            off = (unsigned)(p + 1 - line);
            ri.info_off = off > 0xFFFFu ? 0xFFFFu : (uint16_t)off;
            if (!malformed) {
                ++p;
                if (p[0] == '.' && (p[1] == '\t' || p[1] == '\n')) {
                    ++p;
                } else {
                    for (;;) {
                        const uint8_t* key = p;
                        unsigned hash = 2166136261u, c = *p;
                        while (c != '=' && c != ';' && c != '\t' && c != '\n') {
                            hash = (hash ^ c) * 16777619u;
                            c = *++p;
                        }
                        const int t = find_tag(sp, key, (int)(p - key), hash);
                        const unsigned kind = t >= 0 ? s_tags[t].info_kind : 0u;
                        if (c == '=') {
                            ++p;
                            if (kind)
                                parse_value(sp, s_tags[t], kind, p, ';');
                            p = skip_to(p, ';');
                        } else if (kind) {  // key without a value: typed None / ()
                            for (int s = s_tags[t].first_slot; s < s_tags[t].first_slot + s_tags[t].n_slots; ++s)
                                store_slot(sp, s, RAW_MISSING);
                        }
                        if (*p == ';') {
                            ++p;
                            continue;
                        }
                        break;
                    }
                }
            }
            // ---- FORMAT + first sample (FORMAT values override INFO values of the same
            //      name: the reference builds its per-record dict from info.items() +
            //      samples[0].items(), vcftools.py:69-86)
            off = (unsigned)(p + 1 - line);
            ri.format_off = off > 0xFFFFu ? 0xFFFFu : (uint16_t)off;
            if (!malformed && *p == '\t') {
                const uint8_t* fk = p + 1;
                const uint8_t* sv = skip_to(fk, '\t');
                bool have_sample = (*sv == '\t');
                if (have_sample) ++sv;
                if (!(fk[0] == '.' && (fk[1] == '\t' || fk[1] == '\n'))) {
                    for (;;) {
                        const uint8_t* key = fk;
                        unsigned hash = 2166136261u, c = *fk;
                        while (c != ':' && c != '\t' && c != '\n') {
                            hash = (hash ^ c) * 16777619u;
                            c = *++fk;
                        }
                        const int t = find_tag(sp, key, (int)(fk - key), hash);
                        const unsigned kind = t >= 0 ? s_tags[t].fmt_kind : 0u;
                        if (have_sample) {
                            if (kind)
                                parse_value(sp, s_tags[t], kind, sv, ':');
                            sv = skip_to(sv, ':');
                            if (*sv == ':')
                                ++sv;
                            else
                                have_sample = false;
                        } else if (kind) {  // trailing sub-fields dropped: missing
                            for (int s = s_tags[t].first_slot; s < s_tags[t].first_slot + s_tags[t].n_slots; ++s)
                                store_slot(sp, s, RAW_MISSING);
                        }
                        if (*fk == ':') {
                            ++fk;
                            continue;
                        }
                        break;
                    }
                }
            }
            // ---- fixed-column slots
            for (int s = plan.first_fixed_slot; s < n_slots; ++s) {
                switch (s_slots[s].reducer) {
                    case RED_FIX_QUAL: store_slot(sp, s, qual_bits); break;
                    case RED_FIX_ALLELE0: store_slot_f(sp, s, a0); break;
                    case RED_FIX_ALLELE1: store_slot(sp, s, a1_missing ? RAW_MISSING : __float_as_uint(a1)); break;
                    case RED_FIX_INDEL: store_slot_f(sp, s, indel ? 1.f : 0.f); break;
                    case RED_FIX_NALLELES: store_slot_f(sp, s, (float)n_alleles); break;
                    default: break;
                }
            }
            if (malformed) atomicMin(err, ugvc_pack_error(rec, 0xFFFF, REASON_MALFORMED_LINE));
            if (cg) ri.flags |= 1u;
            cg_local += cg ? 1u : 0u;
            if (recinfo) *reinterpret_cast<uint4*>(&recinfo[rec]) = *reinterpret_cast<const uint4*>(&ri);
        }
        __syncthreads();
        // coalesced columnar write-out of the slot tile
        const long long base = tile * K1_TPB;
        const int valid = (int)min((long long)K1_TPB, n_rec - base);
        for (int s = 0; s < n_slots; ++s)
            if (threadIdx.x < valid) raw[(size_t)s * row_stride + base + threadIdx.x] = s_tile[s * K1_TPB + threadIdx.x];
        __syncthreads();
    }
    // CG-insertion counter: warp ballot-free reduction, one atomic per warp
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) cg_local += __shfl_xor_sync(0xffffffffu, cg_local, s);
    if ((threadIdx.x & 31) == 0 && cg_local) atomicAdd((unsigned long long*)&counts[3], (unsigned long long)cg_local);
}

size_t k1_smem_bytes(const DevPlan& plan) {
    size_t b = (size_t)plan.h.n_tags * sizeof(PlanTag);
    b += (size_t)((plan.h.n_slots + 1) & ~1u) * sizeof(PlanSlot);
    b += (size_t)((plan.h.n_dicts + 1) & ~1u) * sizeof(PlanDict);
    b += (size_t)plan.h.n_dict_strings * sizeof(PlanString);
    b += 256;
    b += (size_t)plan.h.n_slots * K1_TPB * 4;
    return b;
}

void launch_k1(const DevPlan& plan, const uint8_t* d_text, const int64_t* line_start, const int64_t* d_n_records,
               uint32_t* raw, size_t row_stride, ugvc_recinfo* recinfo, unsigned long long* d_err,
               long long* d_counts, int sm_count, cudaStream_t st) {
    const size_t smem = k1_smem_bytes(plan);
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    k1_parse<<<sm_count * per_sm, K1_TPB, smem, st>>>(plan, d_text, line_start, d_n_records, raw, row_stride, recinfo,
                                                      d_err, d_counts);
}

// ------------------------------------------------------------------------------------------
// K2: feature assembly
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K2_TPB) k2_features(const __grid_constant__ DevPlan plan,
                                                      const uint32_t* __restrict__ raw, size_t row_stride,
                                                      const int64_t* __restrict__ n_records_p,
                                                      float* __restrict__ feats, unsigned long long* err) {
    __shared__ PlanFeature s_feat[UGVC_MAX_FEATURES];
    const int F = plan.h.n_features;
    for (int i = threadIdx.x; i < F; i += K2_TPB) s_feat[i] = plan.feats[i];
    __syncthreads();
    const long long n_rec = *n_records_p;
    for (long long rec = (long long)blockIdx.x * K2_TPB + threadIdx.x; rec < n_rec;
         rec += (long long)gridDim.x * K2_TPB) {
#pragma unroll 4
        for (int f = 0; f < F; ++f) {
            const PlanFeature pf = s_feat[f];
            const uint32_t bits = __ldg(&raw[(size_t)pf.slot * row_stride + rec]);
            float v = __uint_as_float(bits);
            if (bits == RAW_ABSENT) {
                if (pf.absent_pol == POL_VALUE) v = pf.absent_val;
                else atomicMin(err, ugvc_pack_error(rec, f, REASON_BAD_VALUE));
            } else if (bits == RAW_MISSING) {
                if (pf.missing_pol == POL_VALUE) v = pf.missing_val;
                else atomicMin(err, ugvc_pack_error(rec, f, REASON_NULL_FEATURE));
            } else if (bits == RAW_ERR) {
                atomicMin(err, ugvc_pack_error(rec, f, REASON_BAD_VALUE));
            }
            feats[(size_t)f * row_stride + rec] = v;
        }
        for (unsigned c = 0; c < plan.h.n_checks; ++c) {
            const PlanCheck ck = plan.checks[c];
            const uint32_t bits = __ldg(&raw[(size_t)ck.slot * row_stride + rec]);
            if (bits == RAW_ABSENT || bits == RAW_MISSING) continue;  // handled by the feature policies
            const float v = __uint_as_float(bits);
            const bool ok = ck.kind == 0 ? (v <= ck.bound) : (v >= ck.bound);
            if (!ok) atomicMin(err, ugvc_pack_error(rec, 0xFFFF, ck.kind == 0 ? REASON_TOO_MANY_ELEMS : REASON_BAD_VALUE));
        }
    }
}

void launch_k2(const DevPlan& plan, const uint32_t* raw, size_t row_stride, const int64_t* d_n_records, float* feats,
               unsigned long long* d_err, int sm_count, cudaStream_t st) {
    k2_features<<<sm_count * 8, K2_TPB, 0, st>>>(plan, raw, row_stride, d_n_records, feats, d_err);
}

// ------------------------------------------------------------------------------------------
// K3: inference + score math + FILTER decision
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double expit64(double x) { return 1.0 / (1.0 + exp(-x)); }

template <int CMP>
__device__ __forceinline__ int walk_tree(const PlanNode* __restrict__ nodes, const float* __restrict__ x_col,
                                         int stride) {
    // preorder layout: left child is the next node; returns the leaf row
    int n = 0;
    for (;;) {
        const PlanNode nd = nodes[n];
        if (nd.feature < 0) return __float_as_int(nd.value);
        const float x = x_col[nd.feature * stride];
        const bool left = (CMP == CMP_LE) ? (x <= nd.value) : (x < nd.value);
        n = left ? n + 1 : (int)nd.right;
    }
}

__global__ void __launch_bounds__(K3_TPB) k3_infer(const __grid_constant__ DevPlan plan,
                                                   const float* __restrict__ feats, size_t row_stride,
                                                   const int64_t* __restrict__ n_records_p, double threshold,
                                                   uint8_t* __restrict__ low_score, float* __restrict__ probs,
                                                   double* __restrict__ qual_out, long long* counts) {
    extern __shared__ __align__(16) uint8_t smem3[];
    float* tile = reinterpret_cast<float*>(smem3);  // [F][K3_TPB]
    const int F = plan.h.n_features, K = plan.h.n_classes, O = plan.h.n_outputs;
    const long long n_rec = *n_records_p;
    const long long n_tiles = (n_rec + K3_TPB - 1) / K3_TPB;
    unsigned n_low = 0, n_seen = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long rec = t * K3_TPB + threadIdx.x;
        const bool active = rec < n_rec;
        __syncthreads();
        // stage the feature tile: row f of the column-major matrix is a coalesced 512 B read
        for (int f = 0; f < F; ++f) tile[f * K3_TPB + threadIdx.x] = active ? feats[(size_t)f * row_stride + rec] : 0.f;
        __syncthreads();
        if (!active) continue;
        const float* x = tile + threadIdx.x;
        double p[UGVC_MAX_CLASSES];
        switch (plan.h.model_kind) {
            case MODEL_LOGISTIC: {
                double z[UGVC_MAX_CLASSES];
                for (int o = 0; o < O; ++o) {
                    double acc = 0.0;
                    const double* w = plan.coef + (size_t)o * F;
                    for (int f = 0; f < F; ++f) acc = fma((double)x[f * K3_TPB], __ldg(&w[f]), acc);
                    z[o] = acc + plan.intercept[o];
                }
                if (O == 1) {
                    p[1] = expit64(z[0]);
                    p[0] = 1.0 - p[1];
                } else {
                    double mx = z[0];
                    for (int o = 1; o < O; ++o) mx = fmax(mx, z[o]);
                    double s = 0.0;
                    for (int o = 0; o < O; ++o) {
                        p[o] = exp(z[o] - mx);
                        s += p[o];
                    }
                    for (int o = 0; o < O; ++o) p[o] /= s;
                }
                break;
            }
            case MODEL_GB_SKLEARN: {
                // raw = init + sum_t (learning_rate * leaf_t) in fp64, tree order
                // (sklearn _gradient_boosting.predict_stages); leaves are pre-scaled on the host
                double z[UGVC_MAX_CLASSES];
                for (int o = 0; o < O; ++o) z[o] = plan.h.init[o];
                for (unsigned tr = 0; tr < plan.h.n_trees; ++tr) {
                    const int leaf = walk_tree<CMP_LE>(plan.nodes + plan.tree_root[tr], x, K3_TPB);
                    const int o = plan.tree_out[tr];
                    z[o] = __dadd_rn(z[o], plan.leaves[leaf]);
                }
                if (O == 1) {
                    p[1] = expit64(z[0]);
                    p[0] = 1.0 - p[1];
                } else {
                    // softmax as exp(raw - logsumexp(raw))
                    double mx = z[0];
                    for (int o = 1; o < O; ++o) mx = fmax(mx, z[o]);
                    double s = 0.0;
                    for (int o = 0; o < O; ++o) s += exp(z[o] - mx);
                    const double lse = mx + log(s);
                    for (int o = 0; o < O; ++o) p[o] = exp(z[o] - lse);
                }
                break;
            }
            case MODEL_RF_SKLEARN: {
                // mean over trees of the per-leaf class fractions, fp64, tree order
                for (int k = 0; k < K; ++k) p[k] = 0.0;
                for (unsigned tr = 0; tr < plan.h.n_trees; ++tr) {
                    const int leaf = walk_tree<CMP_LE>(plan.nodes + plan.tree_root[tr], x, K3_TPB);
                    const double* lv = plan.leaves + (size_t)leaf * plan.h.leaf_width;
                    for (int k = 0; k < K; ++k) p[k] = __dadd_rn(p[k], lv[k]);
                }
                for (int k = 0; k < K; ++k) p[k] = p[k] / (double)plan.h.n_trees;
                break;
            }
            case MODEL_XGB: {
                // xgboost CPU predictor: fp32 margin accumulated in tree order, fp32 sigmoid / softmax
                float z[UGVC_MAX_CLASSES];
                for (int o = 0; o < O; ++o) z[o] = (float)plan.h.init[o];
                for (unsigned tr = 0; tr < plan.h.n_trees; ++tr) {
                    const int leaf = walk_tree<CMP_LT>(plan.nodes + plan.tree_root[tr], x, K3_TPB);
                    const int o = plan.tree_out[tr];
                    z[o] = __fadd_rn(z[o], (float)plan.leaves[leaf]);
                }
                if (O == 1) {
                    const float p1 = 1.0f / (1.0f + expf(-z[0]));
                    p[1] = (double)p1;
                    p[0] = (double)(1.0f - p1);
                } else {
                    float mx = z[0];
                    for (int o = 1; o < O; ++o) mx = fmaxf(mx, z[o]);
                    double s = 0.0;
                    float e[UGVC_MAX_CLASSES];
                    for (int o = 0; o < O; ++o) {
                        e[o] = expf(z[o] - mx);
                        s += (double)e[o];
                    }
                    for (int o = 0; o < O; ++o) p[o] = (double)(e[o] / (float)s);
                }
                break;
            }
            default:
                for (int k = 0; k < K; ++k) p[k] = 0.0;
        }
        // phred = -10 log10(lik + 1e-10); qual = clip(30 + ph[0] - min(ph[1:]), 0, inf)   (fp64)
        double ph0 = -10.0 * log10(p[0] + 1e-10);
        double mn = -10.0 * log10(p[1] + 1e-10);
        for (int k = 2; k < K; ++k) mn = fmin(mn, -10.0 * log10(p[k] + 1e-10));
        double q = __dadd_rn(__dadd_rn(30.0, ph0), -mn);
        q = q < 0.0 ? 0.0 : q;
        const bool low = q <= threshold;
        low_score[rec] = low ? 1 : 0;
        qual_out[rec] = q;
        for (int k = 0; k < K; ++k) probs[(size_t)rec * K + k] = (float)p[k];
        n_low += low ? 1u : 0u;
        n_seen += 1u;
    }
    // pass / fail counters: warp shuffle reduction, one atomic per warp
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        n_low += __shfl_xor_sync(0xffffffffu, n_low, s);
        n_seen += __shfl_xor_sync(0xffffffffu, n_seen, s);
    }
    if ((threadIdx.x & 31) == 0 && n_seen) {
        atomicAdd((unsigned long long*)&counts[0], (unsigned long long)n_seen);
        atomicAdd((unsigned long long*)&counts[1], (unsigned long long)n_low);
        atomicAdd((unsigned long long*)&counts[2], (unsigned long long)(n_seen - n_low));
    }
}

size_t k3_smem_bytes(const DevPlan& plan) { return (size_t)plan.h.n_features * K3_TPB * sizeof(float); }

void launch_k3(const DevPlan& plan, const float* feats, size_t row_stride, const int64_t* d_n_records,
               double threshold, uint8_t* low_score, float* probs, double* qual, long long* d_counts,
               int sm_count, cudaStream_t st) {
    const size_t smem = k3_smem_bytes(plan);
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    k3_infer<<<sm_count * per_sm, K3_TPB, smem, st>>>(plan, feats, row_stride, d_n_records, threshold, low_score,
                                                      probs, qual, d_counts);
}

cudaError_t kernels_configure(const DevPlan& plan) {
    cudaError_t e = cudaFuncSetAttribute(k1_parse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)k1_smem_bytes(plan));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k3_infer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_smem_bytes(plan));
}
