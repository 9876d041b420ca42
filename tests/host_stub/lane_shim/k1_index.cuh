// TEST INFRASTRUCTURE ONLY -- host stand-in for sickle_b200/csrc/k1_index.cuh (see sk_device.cuh here).
#pragma once
#include "sk_device.cuh"

namespace sk {
// 4 flag bits (bit i = byte i flagged) from a word of 0x80 flags: bits 7/15/23/31 -> 28/29/30/31.
inline uint32_t flags_to_nibble(uint32_t f) { return (f * 0x00204081u) >> 28; }
}  // namespace sk
