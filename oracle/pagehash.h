// ORACLE (test infrastructure): the hashing half of the emulator's segment splitter, restated on the CPU.
//   poseidon(bytes)        reference emulator/src/memory.rs:43-79   (rate 8 x u32 little endian, overwrite absorb, pad10*1,
//                                                                    digest = the first 4 state words)
//   hash_page              memory.rs:81-89
//   CONST_HASH_PAGES       memory.rs:91-125  (what an untouched hash page of level 0..2 holds)
//   update_page_hash       memory.rs:378-436 (dirty pages -> 32-byte hashes written into the L1 / L2 / root hash pages)
//   compute_image_id       memory.rs:438-471 (registers into the root page, root hash, image id over the word-swapped root
//                                             hash and the pc)
// as split_segment drives them (emulator/src/state.rs:1460-1478).  Pinned by the Poseidon known answers only (the reference
// holds no image-id fixture): "parity unpinned" beyond the permutation, as for the rest of the plonky2 boundary.
#pragma once
#include "poseidon.h"
#include <cstdint>
#include <map>
#include <stdexcept>
#include <vector>

namespace orc {

constexpr size_t PAGE_SIZE = 4096;
constexpr uint32_t MAX_MEMORY = 0x80000000u, ROOT_PAGE = 0x81020u, REGISTERS_OFFSET = 0x400, REGISTERS_BYTES = 39 * 4;
typedef std::array<uint8_t, PAGE_SIZE> Page;
typedef std::array<uint8_t, 32> Hash32;

static inline void poseidon_bytes(const uint8_t* in, size_t l, uint64_t out[4]) {
    const size_t RATE_BYTES = 32, chunks = l / RATE_BYTES + 1;
    std::vector<uint8_t> input(in, in + l);
    input.resize(chunks * RATE_BYTES, 0);
    if (l % RATE_BYTES == RATE_BYTES - 1) input[l] = 0x81;
    else { input[l] = 1; input[chunks * RATE_BYTES - 1] = 0x80; }
    PState st;
    for (auto& x : st) x = Fp(0);
    for (size_t b = 0; b < chunks; b++) {
        for (int i = 0; i < 8; i++) {
            const uint8_t* p = &input[b * RATE_BYTES + 4 * i];
            st[i] = Fp((uint64_t)p[0] | (uint64_t)p[1] << 8 | (uint64_t)p[2] << 16 | (uint64_t)p[3] << 24);
        }
        poseidon_naive(st);
    }
    for (int i = 0; i < 4; i++) out[i] = st[i].v;
}
static inline Hash32 hash_page(const uint8_t* data) {
    uint64_t h[4];
    poseidon_bytes(data, PAGE_SIZE, h);
    Hash32 o;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) o[8 * i + j] = (uint8_t)(h[i] >> (8 * j));
    return o;
}
// level 0: a page of the hash of the zero page repeated; level k: of the hash of level k - 1
static inline const Page& const_hash_page(int level) {
    static std::vector<Page> pages;
    if (pages.empty()) {
        Page zero{};
        Hash32 base = hash_page(zero.data());
        for (int lv = 0; lv < 3; lv++) {
            Page p;
            for (size_t i = 0; i < PAGE_SIZE / 32; i++) std::copy(base.begin(), base.end(), p.begin() + 32 * i);
            pages.push_back(p);
            base = hash_page(p.data());
        }
    }
    return pages.at(level);
}

struct PageTree {
    std::map<uint32_t, Page> hash_pages;            // the pages at and above MAX_MEMORY
    // set_hash_range: returns the index of the hash page that was written
    uint32_t set_hash_range(uint32_t page_index, const Hash32& h, int level) {
        const uint32_t hash_addr = (page_index << 5) + MAX_MEMORY, hp = hash_addr >> 12, off = hash_addr & 0xFFF;
        auto it = hash_pages.find(hp);
        if (it == hash_pages.end()) it = hash_pages.emplace(hp, const_hash_page(level)).first;
        std::copy(h.begin(), h.end(), it->second.begin() + off);
        return hp;
    }
    // update_page_hash over the dirty main-memory pages, then compute_image_id
    void split(const uint32_t* idx, const uint8_t* pages, size_t n, const uint8_t* registers, uint32_t pc, Hash32& image_id, Hash32& root_hash) {
        std::map<uint32_t, int> dirty[3];
        for (size_t k = 0; k < n; k++) dirty[1][set_hash_range(idx[k], hash_page(pages + k * PAGE_SIZE), 0)] = 1;
        for (int level = 1; level <= 2; level++)
            for (auto& kv : dirty[level]) {
                const uint32_t hp = set_hash_range(kv.first, hash_page(hash_pages.at(kv.first).data()), level);
                if (level < 2) dirty[level + 1][hp] = 1;
            }
        auto it = hash_pages.find(ROOT_PAGE);
        if (it == hash_pages.end()) throw std::runtime_error("compute image ID fail");
        std::copy(registers, registers + REGISTERS_BYTES, it->second.begin() + REGISTERS_OFFSET);
        root_hash = hash_page(it->second.data());
        uint8_t fin[36];
        for (int i = 0; i < 32; i += 4) { fin[i] = root_hash[i + 3]; fin[i + 1] = root_hash[i + 2]; fin[i + 2] = root_hash[i + 1]; fin[i + 3] = root_hash[i]; }
        for (int j = 0; j < 4; j++) fin[32 + j] = (uint8_t)(pc >> (8 * j));
        uint64_t h[4];
        poseidon_bytes(fin, 36, h);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) image_id[8 * i + j] = (uint8_t)(h[i] >> (8 * j));
    }
};

}  // namespace orc
