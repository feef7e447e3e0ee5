// hostz.h -- host-side DEFLATE / CRC-32 for the BGZF readers (bam.cu): what the reference gets from htslib's
// bgzf.c + zlib (rust-htslib under lib-coverage/src/lib.rs:533-547).  Pure host code, no device part.
#pragma once
#include <cstddef>
#include <cstdint>

namespace cnvb {

// Bytes that must be readable past the end of the input of inflate_raw_fast (it refills its bit buffer with
// unaligned 8-byte loads and checks the consumed length afterwards).
constexpr size_t kInflateSlack = 16;

// CRC-32 of the gzip / BGZF trailer (polynomial 0xEDB88320) of n bytes: folding by carry-less multiplication
// where the CPU has PCLMULQDQ, zlib's crc32 otherwise.  Same value as zlib's crc32(0, p, n).
uint32_t crc32_fast(const uint8_t *p, size_t n);

// Raw DEFLATE stream (RFC 1951) of in_len bytes -> exactly out_len bytes at out.  Returns false on ANYTHING
// irregular (corrupt or truncated stream, output that is not exactly out_len bytes, input not used up to its
// last byte): the caller then repeats the block with zlib, whose verdict stands.  Never writes outside
// [out, out + out_len) and never reads outside [in, in + in_len + kInflateSlack).
bool inflate_raw_fast(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

}  // namespace cnvb
