// deflate.cuh -- BGZF deflate on the device (RFC 1951 fixed-Huffman blocks inside RFC 1952 gzip members).
//
// the reference writes its output through htslib's BGZF writer
// (filter_variants_pipeline.py:115,228); compressing the spliced records on the GPU means only compressed bytes
// cross PCIe on the way out as well, and the host's zlib -- the slowest stage of the file-to-file tool -- leaves
// the path.  Any DEFLATE stream a reader inflates back to the same bytes is a valid output: this encoder trades
// ratio for simplicity and speed.
//
// One thread per block of DEF_CHUNK input bytes (self-contained, like every BGZF block): greedy LZ77 with a
// 2048-entry hash of the last position of every 4-byte prefix (thread-private table in SHARED memory: in local
// memory every probe was an L2 round trip and a 5 M-record file took 1.1 s to encode), matches of 4..258 bytes at
// distances up to 32 KiB, fixed Huffman codes (no code-length header to build), a 64-bit bit buffer flushed with
// 32-bit stores.  The CRC32 of the block (gzip footer; htslib verifies it) is computed by the same thread,
// slicing by four bytes.  A block that would not shrink is stored.  Output: a complete BGZF block (18-byte header,
// payload, CRC32, ISIZE) at out, its size returned; the caller packs the blocks (prefix sum of the sizes).
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

#define DEF_CHUNK 57344u        // input bytes per BGZF block: 9 bits per literal at worst stays under 64 KiB
#define DEF_BLOCK_STRIDE 65536u // bytes reserved per output block
#define DEF_HASH_BITS 11     // 2048 entries x 2 bytes: the tables of a 32-thread CTA fill 128 KiB of shared memory
#define DEF_MIN_MATCH 4u
#define DEF_MAX_MATCH 258u

struct DefTables {
    uint32_t crc[4][256];   // slice-by-4 CRC32 tables
    uint16_t len_code[259]; // length 3..258 -> symbol - 257 in the low 5 bits, extra-bit count in bits 8..10
    uint16_t len_base[29];
    uint8_t dist_code[512]; // zlib's d_code table
    uint16_t dist_base[30];
    uint8_t dist_extra[30];
    uint32_t x2n[32];       // x^(2^k) mod the CRC polynomial (reflected): the CRC32 of a concatenation from its parts
};

// a * b mod the CRC-32 polynomial, both in the reflected representation (bit 31 = x^0)
__host__ __device__ inline uint32_t def_multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 31; i >= 0; --i) {
        p ^= b & (0u - ((a >> i) & 1u));
        b = (b >> 1) ^ (0xEDB88320u & (0u - (b & 1u)));
    }
    return p;
}

// host: fill the tables (uploaded once per context)
static inline void def_build_tables(DefTables& t) {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
        t.crc[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i) {
        t.crc[1][i] = (t.crc[0][i] >> 8) ^ t.crc[0][t.crc[0][i] & 0xFF];
        t.crc[2][i] = (t.crc[1][i] >> 8) ^ t.crc[0][t.crc[1][i] & 0xFF];
        t.crc[3][i] = (t.crc[2][i] >> 8) ^ t.crc[0][t.crc[2][i] & 0xFF];
    }
    t.x2n[0] = 0x40000000u;  // x^1
    for (int k = 1; k < 32; ++k) t.x2n[k] = def_multmodp(t.x2n[k - 1], t.x2n[k - 1]);
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    uint32_t len = 3;
    for (int c = 0; c < 29; ++c) {
        t.len_base[c] = (uint16_t)len;
        const uint32_t n = c == 28 ? 1u : (1u << lext[c]);
        for (uint32_t k = 0; k < n && len <= 258; ++k, ++len) t.len_code[len] = (uint16_t)(c | (lext[c] << 8));
        if (c == 27) len = 258;  // code 284 stops at 257, code 285 is length 258 alone
    }
    t.len_code[258] = (uint16_t)28;
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    uint32_t dist = 1;
    for (int c = 0; c < 30; ++c) {
        t.dist_base[c] = (uint16_t)dist;
        t.dist_extra[c] = dext[c];
        for (uint32_t k = 0; k < (1u << dext[c]); ++k, ++dist) {
            const uint32_t d0 = dist - 1;  // 0-based
            if (d0 < 256) t.dist_code[d0] = (uint8_t)c;
            else t.dist_code[256 + (d0 >> 7)] = (uint8_t)c;
        }
    }
}

struct DefBits {
    uint8_t* out;   // 4-byte aligned
    uint32_t pos;   // bytes written
    uint64_t buf;
    uint32_t cnt;
    __host__ __device__ inline void put(uint32_t bits, uint32_t n) {  // n <= 32, LSB first
        buf |= (uint64_t)bits << cnt;
        cnt += n;
        if (cnt >= 32) {
            *reinterpret_cast<uint32_t*>(out + pos) = (uint32_t)buf;
            pos += 4;
            buf >>= 32;
            cnt -= 32;
        }
    }
    __host__ __device__ inline void finish() {  // pad to a byte boundary, flush
        while (cnt > 0) {
            out[pos++] = (uint8_t)buf;
            buf >>= 8;
            cnt = cnt > 8 ? cnt - 8 : 0;
        }
    }
};

__host__ __device__ inline uint32_t def_rev(uint32_t code, uint32_t n) {  // Huffman codes go out most significant bit first
#if defined(__CUDA_ARCH__)
    return __brev(code) >> (32u - n);
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; ++i) r |= ((code >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}
__host__ __device__ inline void def_put_litlen(DefBits& b, uint32_t sym) {
    if (sym < 144) b.put(def_rev(0x30 + sym, 8), 8);
    else if (sym < 256) b.put(def_rev(0x190 + (sym - 144), 9), 9);
    else if (sym < 280) b.put(def_rev(sym - 256, 7), 7);
    else b.put(def_rev(0xC0 + (sym - 280), 8), 8);
}
__host__ __device__ inline uint32_t def_ld4(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    return sh ? (w[0] >> sh) | (w[1] << (32u - sh)) : w[0];
}

// One BGZF block from in[0, n) (n <= DEF_CHUNK; 4 readable bytes after in + n).  out: 4-byte aligned, DEF_BLOCK_STRIDE
// bytes.  head: 1 << DEF_HASH_BITS entries of scratch.  Returns the block size.
__host__ __device__ inline uint32_t def_block(const uint8_t* in, uint32_t n, uint8_t* out, uint16_t* head, const DefTables& T) {
    // ---- gzip / BGZF header (BSIZE patched at the end)
    const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
    for (int i = 0; i < 18; ++i) out[i] = hdr[i];
    // ---- CRC32, four bytes a step
    uint32_t crc = 0xFFFFFFFFu;
    {
        uint32_t i = 0;
        for (; i + 4 <= n; i += 4) {
            const uint32_t w = def_ld4(in + i) ^ crc;
            crc = T.crc[3][w & 0xFF] ^ T.crc[2][(w >> 8) & 0xFF] ^ T.crc[1][(w >> 16) & 0xFF] ^ T.crc[0][w >> 24];
        }
        for (; i < n; ++i) crc = (crc >> 8) ^ T.crc[0][(crc ^ in[i]) & 0xFF];
        crc = ~crc;
    }
    // ---- fixed-Huffman DEFLATE block
    DefBits b;
    b.out = out + 16;  // the payload starts at 18: the bit buffer begins with the two BSIZE bytes (patched at the end), so
    b.pos = 0;         // the 32-bit stores stay aligned and nothing has to be moved afterwards
    b.buf = 0;
    b.cnt = 16;
    b.put(1u, 1);  // BFINAL
    b.put(1u, 2);  // BTYPE = 01
    for (uint32_t i = 0; i < (1u << DEF_HASH_BITS); ++i) head[i] = 0xFFFFu;
    uint32_t i = 0;
    while (i < n) {
        uint32_t best = 0, dist = 0;
        if (i + DEF_MIN_MATCH <= n) {
            const uint32_t w = def_ld4(in + i);
            const uint32_t h = (w * 2654435761u) >> (32 - DEF_HASH_BITS);
            const uint32_t cand = head[h];
            head[h] = (uint16_t)i;
            if (cand != 0xFFFFu && i - cand <= 32768u && def_ld4(in + cand) == w) {
                uint32_t len = 4;
                const uint32_t maxlen = n - i < DEF_MAX_MATCH ? n - i : DEF_MAX_MATCH;
                while (len + 4 <= maxlen) {
                    const uint32_t x = def_ld4(in + cand + len) ^ def_ld4(in + i + len);
                    if (x) {
#if defined(__CUDA_ARCH__)
                        len += (uint32_t)(__ffs((int)x) - 1) >> 3;
#else
                        len += (uint32_t)__builtin_ctz(x) >> 3;
#endif
                        break;
                    }
                    len += 4;
                }
                if (len + 4 > maxlen)
                    while (len < maxlen && in[cand + len] == in[i + len]) ++len;
                best = len > maxlen ? maxlen : len;
                dist = i - cand;
            }
        }
        if (best >= DEF_MIN_MATCH) {
            const uint32_t lc = T.len_code[best], lsym = lc & 31u, lx = lc >> 8;
            def_put_litlen(b, 257u + lsym);
            if (lx) b.put(best - T.len_base[lsym], lx);
            const uint32_t d0 = dist - 1;
            const uint32_t dc = d0 < 256 ? T.dist_code[d0] : T.dist_code[256 + (d0 >> 7)];
            b.put(def_rev(dc, 5), 5);
            if (T.dist_extra[dc]) b.put(dist - T.dist_base[dc], T.dist_extra[dc]);
            i += best;
        } else {
            def_put_litlen(b, in[i]);
            ++i;
        }
        if (b.pos + 16 > DEF_BLOCK_STRIDE - 64) break;  // cannot happen for n <= DEF_CHUNK (9 bits per byte at worst)
    }
    def_put_litlen(b, 256u);  // end of block
    b.finish();
    uint32_t payload = b.pos - 2;
    if (payload >= n + 5 || i < n) {
        // did not shrink: one stored block (BFINAL=1, BTYPE=00, LEN, NLEN, bytes)
        uint8_t* p = out + 18;
        p[0] = 1;
        p[1] = (uint8_t)n;
        p[2] = (uint8_t)(n >> 8);
        p[3] = (uint8_t)~n;
        p[4] = (uint8_t)(~n >> 8);
        for (uint32_t k = 0; k < n; ++k) p[5 + k] = in[k];
        payload = n + 5;
    }
    uint8_t* ft = out + 18 + payload;
    ft[0] = (uint8_t)crc;
    ft[1] = (uint8_t)(crc >> 8);
    ft[2] = (uint8_t)(crc >> 16);
    ft[3] = (uint8_t)(crc >> 24);
    ft[4] = (uint8_t)n;
    ft[5] = (uint8_t)(n >> 8);
    ft[6] = (uint8_t)(n >> 16);
    ft[7] = (uint8_t)(n >> 24);
    const uint32_t bsize = 18 + payload + 8;
    out[16] = (uint8_t)(bsize - 1);
    out[17] = (uint8_t)((bsize - 1) >> 8);
    return bsize;
}

// ------------------------------------------------------------------------------------------
// pieces of the warp-per-block encoder (fileio.cu: fio_deflate_warp), shared with its host model
// (hostio.cpp: ugvc_test_deflate_block_lanes) so that everything but the shuffles is exercised on the CPU
// ------------------------------------------------------------------------------------------
// CRC32 (the usual pre / post inversion) of p[0, n)
__host__ __device__ inline uint32_t def_crc_slice(const DefTables& T, const uint8_t* p, uint32_t n) {
    uint32_t crc = 0xFFFFFFFFu, i = 0;
    for (; i + 4 <= n; i += 4) {
        const uint32_t w = def_ld4(p + i) ^ crc;
        crc = T.crc[3][w & 0xFF] ^ T.crc[2][(w >> 8) & 0xFF] ^ T.crc[1][(w >> 16) & 0xFF] ^ T.crc[0][w >> 24];
    }
    for (; i < n; ++i) crc = (crc >> 8) ^ T.crc[0][(crc ^ p[i]) & 0xFF];
    return ~crc;
}
// crc(A || B) from crc(A), crc(B) and |B|: crc(A) * x^(8 |B|) + crc(B)
__host__ __device__ inline uint32_t def_crc_combine(const DefTables& T, uint32_t crc_a, uint32_t crc_b, uint32_t len_b) {
    uint32_t p = 0x80000000u;  // x^0
    for (uint32_t n = len_b, k = 3; n; n >>= 1, ++k)
        if (n & 1u) p = def_multmodp(T.x2n[k & 31u], p);
    return def_multmodp(p, crc_a) ^ crc_b;
}
// the bits of one token, LSB first: a literal (len == 0) or a match of len bytes dist back; at most 31 bits
__host__ __device__ inline uint32_t def_token(const DefTables& T, uint32_t lit, uint32_t len, uint32_t dist, uint32_t& nbits) {
    if (len == 0) {
        if (lit < 144) {
            nbits = 8;
            return def_rev(0x30 + lit, 8);
        }
        nbits = 9;
        return def_rev(0x190 + (lit - 144), 9);
    }
    const uint32_t lc = T.len_code[len], lsym = lc & 31u, lx = lc >> 8;
    uint32_t bits, n;
    if (lsym < 23) {  // symbols 257..279: seven bits
        bits = def_rev(lsym + 1, 7);
        n = 7;
    } else {
        bits = def_rev(0xC0 + (lsym - 23), 8);
        n = 8;
    }
    if (lx) {
        bits |= (len - T.len_base[lsym]) << n;
        n += lx;
    }
    const uint32_t d0 = dist - 1;
    const uint32_t dc = d0 < 256 ? T.dist_code[d0] : T.dist_code[256 + (d0 >> 7)];
    bits |= def_rev(dc, 5) << n;
    n += 5;
    const uint32_t dx = T.dist_extra[dc];
    if (dx) {
        bits |= (dist - T.dist_base[dc]) << n;
        n += dx;
    }
    nbits = n;
    return bits;
}
// the longest match of in[pos..] against in[cand..] given that the first four bytes agree; at most maxlen
__host__ __device__ inline uint32_t def_extend(const uint8_t* in, uint32_t cand, uint32_t pos, uint32_t maxlen) {
    uint32_t l = 4;
    while (l + 4 <= maxlen) {
        const uint32_t x = def_ld4(in + cand + l) ^ def_ld4(in + pos + l);
        if (x) {
#if defined(__CUDA_ARCH__)
            return l + ((uint32_t)(__ffs((int)x) - 1) >> 3);
#else
            return l + ((uint32_t)__builtin_ctz(x) >> 3);
#endif
        }
        l += 4;
    }
    while (l < maxlen && in[cand + l] == in[pos + l]) ++l;
    return l;
}
#define DEFW_HASH(w) (((w) * 2654435761u) >> (32 - DEF_HASH_BITS))
