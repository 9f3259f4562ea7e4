// A DEFLATE (RFC 1951) decoder for the BGZF members of a BAM: the inflate that bounds BAM -> evidence on the host side.
#pragma once
#include <cstddef>
#include <cstdint>

namespace brq {

// Inflates a raw DEFLATE stream of src_len bytes into exactly dst_len bytes.  false = the stream is malformed, does not fill
// dst exactly, or uses something this decoder does not take: dst is then unspecified and the caller falls back to zlib.
// Never reads outside [src, src + src_len) or writes outside [dst, dst + dst_len).
bool fast_inflate(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len);

}  // namespace brq
