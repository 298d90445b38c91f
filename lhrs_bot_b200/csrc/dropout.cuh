// LoRA input dropout (peft lora.Linear: result += lora_B(lora_A(dropout(x))) * scaling, text_modal.py:136-143 configures
// lora_dropout = 0.05 in every shipped yaml).  The mask is a pure function of (call seed, LoRA module index, row, column), so
// the forward T product, the backward dA product and the dX correction regenerate the same bits without storing a mask:
//     key   = mix32(seed_lo ^ mix32(seed_hi + module * 0x9E3779B9))            module = layer * 7 + {q,k,v,o,gate,up,down}
//     word  = mix32(key + ((row >> 1) * (ld >> 1) + (col >> 1)) * 0x9E3779B1)   one word per 2 x 2 block of the input, uint32 wrap
//     byte  = (word >> 8 * (2 * (row & 1) + (col & 1))) & 255
//     keep  = byte >= T,  T = round(p * 256)   ->  drop probability T / 256 (13 / 256 = 0.0508 for the yamls' 0.05), survivors are
//                                                  scaled by 256 / (256 - T)
// mix32 is the 32-bit murmur3 finaliser.  torch's own dropout stream cannot be reproduced (it depends on the launch geometry of
// torch's kernel); the oracle restates THIS function (oracle/llama.py: lora_dropout_mask) and is compared bit for bit.
#pragma once
#include <stdint.h>

namespace lhrs {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t drop_key(uint64_t seed, uint32_t module) {
    return mix32(static_cast<uint32_t>(seed) ^ mix32(static_cast<uint32_t>(seed >> 32) + module * 0x9E3779B9u));
}
__host__ __device__ __forceinline__ int drop_threshold(float p) {
    int t = static_cast<int>(p * 256.0f + 0.5f);
    return t < 0 ? 0 : (t > 255 ? 255 : t);
}
__host__ __device__ __forceinline__ float drop_inv_keep(int t) { return 256.0f / static_cast<float>(256 - t); }
// the 4 draws of the 2 x 2 block that holds (row, col); ld = columns of the dropped-out matrix (even)
__host__ __device__ __forceinline__ uint32_t drop_word(uint32_t key, uint32_t row, uint32_t col, uint32_t ld) {
    return mix32(key + ((row >> 1) * (ld >> 1) + (col >> 1)) * 0x9E3779B1u);
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t word, uint32_t row, uint32_t col, int t) {
    return static_cast<int>((word >> (8u * (2u * (row & 1u) + (col & 1u)))) & 255u) >= t;
}

}  // namespace lhrs
