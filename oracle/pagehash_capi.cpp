// ORACLE (test infrastructure): C entry points over oracle/pagehash.h (the emulator's page hashing restated on the CPU).
#include "pagehash.h"
#include "par.h"
#include <cstring>

using namespace orc;

extern "C" {

// hash_page over n pages (4096 B each) -> n x 32 B
void orc_hash_pages(const uint8_t* pages, size_t n, uint8_t* out) {
    parallel_for(n, [&](size_t k) {
        Hash32 h = hash_page(pages + k * PAGE_SIZE);
        memcpy(out + 32 * k, h.data(), 32);
    });
}
void orc_poseidon_bytes(const uint8_t* in, size_t len, uint64_t* out4) { poseidon_bytes(in, len, out4); }
void orc_const_hash_page(int level, uint8_t* out) { memcpy(out, const_hash_page(level).data(), PAGE_SIZE); }

void* orc_pagetree_create() { return new PageTree(); }
void orc_pagetree_destroy(void* t) { delete (PageTree*)t; }
// update_page_hash over the dirty pages + compute_image_id; returns 0, or -1 when the root page does not exist
int orc_pagetree_split(void* t, const uint32_t* idx, const uint8_t* pages, size_t n, const uint8_t* registers, uint32_t pc,
                       uint8_t* image_id, uint8_t* root_hash) {
    try {
        Hash32 id, root;
        ((PageTree*)t)->split(idx, pages, n, registers, pc, id, root);
        memcpy(image_id, id.data(), 32); memcpy(root_hash, root.data(), 32);
        return 0;
    } catch (const std::exception&) { return -1; }
}
int orc_pagetree_page(const void* t, uint32_t page_index, uint8_t* out) {
    auto& m = ((const PageTree*)t)->hash_pages;
    auto it = m.find(page_index);
    if (it == m.end()) return 0;
    memcpy(out, it->second.data(), PAGE_SIZE);
    return 1;
}
size_t orc_pagetree_count(const void* t) { return ((const PageTree*)t)->hash_pages.size(); }

}  // extern "C"
