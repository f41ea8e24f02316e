// taub_tiff.cu -- host-side decoders for the two byte-oriented TIFF codecs (TIFF 6.0 sections 9 and 13) behind
// taufactor_b200.io.imread: the reference's users load their volumes with tifffile.imread (README.md:51-54) and
// a 512^3 LZW stack costs tens of seconds in an interpreter loop -- two orders of magnitude more than the solve.
// Plain host C++ (no CUDA calls); Deflate stays with zlib on the Python side.
#include <string.h>

#include "taub_common.cuh"

extern "C" {

// PackBits: n in 0..127 copies n+1 literal bytes, n in 129..255 repeats the next byte 257-n times, 128 is a no-op.
// Returns the number of bytes written (decoding stops when dst is full or src is exhausted), < 0 on bad arguments.
int64_t taub_unpackbits(const uint8_t *src, size_t n_src, uint8_t *dst, size_t n_dst)
{
    TAUB_REQUIRE((src || n_src == 0) && (dst || n_dst == 0), "taub_unpackbits: null pointer");
    size_t i = 0, o = 0;
    while (i < n_src && o < n_dst) {
        const unsigned c = src[i++];
        if (c < 128) {
            size_t len = c + 1;
            if (len > n_src - i) len = n_src - i;
            if (len > n_dst - o) len = n_dst - o;
            memcpy(dst + o, src + i, len);
            i += c + 1;
            o += len;
        } else if (c > 128) {
            if (i >= n_src) break;
            size_t len = 257 - c;
            if (len > n_dst - o) len = n_dst - o;
            memset(dst + o, src[i++], len);
            o += len;
        }
    }
    return (int64_t)o;
}

// TIFF LZW: MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, "early change" (the width grows
// one code before the table is full).  The string table is kept as (prefix code, last byte, length) and strings are
// written back to front.  Returns the number of bytes written, -1 for a corrupt stream (message in
// taub_last_error()).
int64_t taub_unlzw(const uint8_t *src, size_t n_src, uint8_t *dst, size_t n_dst)
{
    TAUB_REQUIRE((src || n_src == 0) && (dst || n_dst == 0), "taub_unlzw: null pointer");
    enum { CLEAR = 256, EOI = 257, FIRST = 258, MAXC = 4096 };
    static thread_local uint16_t prefix[MAXC];
    static thread_local uint8_t last[MAXC], first[MAXC];
    static thread_local uint32_t length[MAXC];
    for (int c = 0; c < 256; ++c) {
        prefix[c] = 0;
        last[c] = first[c] = (uint8_t)c;
        length[c] = 1;
    }
    uint32_t bitbuf = 0;
    int nbits = 0, width = 9, next = FIRST, prev = -1;
    size_t i = 0, o = 0;
    while (o < n_dst) {
        while (nbits < width && i < n_src) {
            bitbuf = (bitbuf << 8) | src[i++];
            nbits += 8;
        }
        if (nbits < width) break;
        const int code = (int)((bitbuf >> (nbits - width)) & ((1u << width) - 1u));
        nbits -= width;
        if (code == EOI) break;
        if (code == CLEAR) {
            next = FIRST;
            width = 9;
            prev = -1;
            continue;
        }
        int emit;   // the code whose string is written now
        if (prev < 0) {
            TAUB_REQUIRE(code < 256, "taub_unlzw: corrupt stream (first code %d after a clear)", code);
            emit = code;
        } else {
            if (next < MAXC) {      // new entry = string(prev) + first byte of string(code)  (code == next: of string(prev))
                TAUB_REQUIRE(code <= next, "taub_unlzw: corrupt stream (code %d, table %d)", code, next);
                prefix[next] = (uint16_t)prev;
                first[next] = first[prev];
                last[next] = (code < next) ? first[code] : first[prev];
                length[next] = length[prev] + 1;
                ++next;
            } else {                // table full: nothing is added until the next ClearCode
                TAUB_REQUIRE(code < next, "taub_unlzw: corrupt stream (code %d, table full)", code);
            }
            emit = code;
        }
        // write string(emit) back to front, clipped to the room left
        const size_t len = length[emit];
        size_t room = n_dst - o, skip = len > room ? len - room : 0;
        int c = emit;
        for (size_t k = len; k-- > 0;) {
            if (k < len - skip) dst[o + k] = last[c];
            c = prefix[c];
        }
        o += len - skip;
        prev = emit;
        if (next >= (1 << width) - 1 && width < 12) ++width;
    }
    return (int64_t)o;
}

}  // extern "C"
