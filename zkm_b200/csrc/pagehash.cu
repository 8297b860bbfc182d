// The hashing half of the emulator's segment splitter on the device.  At every segment boundary the
// reference's split_segment (emulator/src/state.rs:1460-1478) re-hashes every memory page the segment wrote
// (Memory::update_page_hash, emulator/src/memory.rs:415-436: 129 Poseidon permutations per 4 KiB page, then the L1 / L2 hash
// pages those digests land in) and derives the image id (compute_image_id, :438-471).  Pages are independent, so the level-0
// pass is one launch over all dirty pages; the two upper levels are a handful of pages each.
//   page_hash_kernel        hash_page = poseidon(bytes) of memory.rs:43-89: rate 8 x u32 (little endian), every block OVERWRITES
//                           the rate, pad10*1 -- a 4096-byte page is 128 data blocks and the block (1, 0, .., 0, 0x80000000)
//   PageTreeDev             the hash pages at and above MAX_MEMORY (host mirror), set_hash_range / alloc_hash_page (:378-413)
// One thread walks one page (the sponge is sequential); 32-thread CTAs spread the warps over the SMs.
#include "dev.cuh"
#include "poseidon_v2.cuh"
#include "poseidon_host.h"
#include "batch.cuh"
#include <array>
#include <cstring>
#include <map>

namespace zkm {

constexpr size_t PAGE_BYTES = 4096, PAGE_WORDS64 = PAGE_BYTES / 8;
constexpr u32 PT_MAX_MEMORY = 0x80000000u, PT_ROOT_PAGE = 0x81020u, PT_REGISTERS_OFFSET = 0x400, PT_REGISTERS_BYTES = 39 * 4;

__global__ void __launch_bounds__(32) page_hash_kernel(const uint4* __restrict__ pages, size_t n_pages, u64* __restrict__ digests) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pages) return;
    const uint4* p = pages + k * (PAGE_BYTES / 16);
    u64 st[12];
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = 0;
#pragma unroll 1
    for (int b = 0; b < (int)(PAGE_BYTES / 32); b++) {
        const uint4 lo = p[2 * b], hi = p[2 * b + 1];
        st[0] = lo.x; st[1] = lo.y; st[2] = lo.z; st[3] = lo.w; st[4] = hi.x; st[5] = hi.y; st[6] = hi.z; st[7] = hi.w;
        poseidon_permute_dev(st);
    }
    st[0] = 1; st[1] = st[2] = st[3] = st[4] = st[5] = st[6] = 0; st[7] = 0x80000000u;      // pad10*1, len % 32 == 0
    poseidon_permute_dev(st);
#pragma unroll
    for (int i = 0; i < 4; i++) digests[4 * k + i] = st[i];
}

// hash_page over n contiguous host pages -> n x 32 bytes (4 little-endian u64 words each)
void hash_pages_dev(const uint8_t* h_pages, size_t n, uint8_t* h_out, cudaStream_t s) {
    if (!n) return;
    ZKM_CHECK(n <= ((size_t)1 << 20), "too many pages in one call");
    DevBuf pages(n * PAGE_WORDS64, s), dig(4 * n, s);
    ZKM_CUDA(cudaMemcpyAsync(pages.p, h_pages, n * PAGE_BYTES, cudaMemcpyHostToDevice, s));
    {
        ProfScope ps("page_hash", s, (double)n * (PAGE_BYTES + 32), (double)n * 129);
        page_hash_kernel<<<(unsigned)((n + 31) / 32), 32, 0, s>>>((const uint4*)pages.p, n, dig.p);
        ZKM_LAUNCHED();
    }
    std::vector<u64> d(4 * n);
    dig.download(d.data(), 4 * n);
    memcpy(h_out, d.data(), 32 * n);                  // little-endian host: the u64 words are the reference's to_le_bytes
}

// poseidon(bytes) on the host for the 36-byte image-id preimage (2 permutations)
static void poseidon_bytes_host(const uint8_t* in, size_t l, uint8_t out[32]) {
    const size_t chunks = l / 32 + 1;
    std::vector<uint8_t> input(chunks * 32, 0);
    memcpy(input.data(), in, l);
    if (l % 32 == 31) input[l] = 0x81; else { input[l] = 1; input[chunks * 32 - 1] = 0x80; }
    u64 st[12] = {0};
    for (size_t b = 0; b < chunks; b++) {
        for (int i = 0; i < 8; i++) { u32 w; memcpy(&w, &input[32 * b + 4 * i], 4); st[i] = w; }
        poseidon_permute_host(st);
    }
    memcpy(out, st, 32);
}

struct PageTreeDev {
    typedef std::array<uint8_t, PAGE_BYTES> Page;
    std::map<u32, Page> hash_pages;
    Page const_pages[3];
    bool have_const = false;
    void ensure_const(cudaStream_t s) {                 // CONST_HASH_PAGES, memory.rs:91-125
        if (have_const) return;
        Page cur{};
        for (int lv = 0; lv < 3; lv++) {
            uint8_t h[32];
            hash_pages_dev(cur.data(), 1, h, s);
            for (size_t i = 0; i < PAGE_BYTES / 32; i++) memcpy(const_pages[lv].data() + 32 * i, h, 32);
            cur = const_pages[lv];
        }
        have_const = true;
    }
    u32 set_hash_range(u32 page_index, const uint8_t* h, int level) {
        const u32 hash_addr = (page_index << 5) + PT_MAX_MEMORY, hp = hash_addr >> 12, off = hash_addr & 0xFFF;
        auto it = hash_pages.find(hp);
        if (it == hash_pages.end()) it = hash_pages.emplace(hp, const_pages[level]).first;
        memcpy(it->second.data() + off, h, 32);
        return hp;
    }
    void split(const u32* idx, const uint8_t* pages, size_t n, const uint8_t* registers, u32 pc, uint8_t* image_id, uint8_t* root_hash,
               cudaStream_t s) {
        ensure_const(s);
        for (size_t k = 0; k < n; k++) ZKM_CHECK(idx[k] < (PT_MAX_MEMORY >> 12), "dirty page index is not a main-memory page");
        std::vector<uint8_t> dig(32 * n);
        hash_pages_dev(pages, n, dig.data(), s);
        std::map<u32, int> dirty;
        for (size_t k = 0; k < n; k++) dirty[set_hash_range(idx[k], &dig[32 * k], 0)] = 1;
        for (int level = 1; level <= 2; level++) {
            std::vector<u32> ids;
            std::vector<uint8_t> buf;
            for (auto& kv : dirty) { ids.push_back(kv.first); const Page& p = hash_pages.at(kv.first); buf.insert(buf.end(), p.begin(), p.end()); }
            dig.assign(32 * ids.size(), 0);
            hash_pages_dev(buf.data(), ids.size(), dig.data(), s);
            dirty.clear();
            for (size_t k = 0; k < ids.size(); k++) dirty[set_hash_range(ids[k], &dig[32 * k], level)] = 1;
        }
        auto it = hash_pages.find(PT_ROOT_PAGE);
        ZKM_CHECK(it != hash_pages.end(), "compute image ID fail");            // memory.rs:443
        memcpy(it->second.data() + PT_REGISTERS_OFFSET, registers, PT_REGISTERS_BYTES);
        hash_pages_dev(it->second.data(), 1, root_hash, s);
        uint8_t fin[36];
        for (int i = 0; i < 32; i += 4) { fin[i] = root_hash[i + 3]; fin[i + 1] = root_hash[i + 2]; fin[i + 2] = root_hash[i + 1]; fin[i + 3] = root_hash[i]; }
        memcpy(fin + 32, &pc, 4);
        poseidon_bytes_host(fin, 36, image_id);
    }
};

}  // namespace zkm

// ---- C ABI (include/zkm_b200.h, "emulator segment splitter") -------------------------------------------------------
#include "../../include/zkm_b200.h"
#include <cstdlib>
#include <string>
#include <vector>

static int pagehash_fail(char** err, const std::exception& e) {
    if (err) {
        const size_t n = strlen(e.what());
        char* m = (char*)malloc(n + 1);
        if (m) memcpy(m, e.what(), n + 1);
        *err = m;
    }
    return -1;
}
#define ZKM_API_BEGIN if (err) *err = nullptr; try {
#define ZKM_API_END } catch (const std::exception& e) { return pagehash_fail(err, e); } return 0;

struct zkm_pagetree { zkm::PageTreeDev t; };

extern "C" {

int zkm_b200_hash_pages(const uint8_t* pages, size_t n_pages, uint8_t* digests_out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK((pages && digests_out) || n_pages == 0, "null argument");
    zkm::hash_pages_dev(pages, n_pages, digests_out, zkm::ctx().stream);
    ZKM_API_END
}
int zkm_b200_pagetree_create(zkm_pagetree_t** out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(out, "null argument");
    *out = new zkm_pagetree;
    ZKM_API_END
}
void zkm_b200_pagetree_destroy(zkm_pagetree_t* t) { delete t; }
int zkm_b200_pagetree_split(zkm_pagetree_t* t, const uint32_t* page_indices, const uint8_t* pages, size_t n_pages,
                            const uint8_t* registers, uint32_t pc, uint8_t* image_id_out, uint8_t* page_hash_root_out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(t && registers && image_id_out && page_hash_root_out && ((page_indices && pages) || n_pages == 0), "null argument");
    t->t.split(page_indices, pages, n_pages, registers, pc, image_id_out, page_hash_root_out, zkm::ctx().stream);
    ZKM_API_END
}
int zkm_b200_pagetree_page(const zkm_pagetree_t* t, uint32_t page_index, uint8_t* out, int* present, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(t && out && present, "null argument");
    auto it = t->t.hash_pages.find(page_index);
    *present = it != t->t.hash_pages.end();
    if (*present) memcpy(out, it->second.data(), zkm::PAGE_BYTES);
    ZKM_API_END
}

// A state resumed from a segment file (State::load_seg, emulator/src/state.rs:141-190; split_seg_into_segs, utils.rs:62-109) brings
// its hash pages in the memory image: they seed the tree before the first split.  Host-only.
int zkm_b200_pagetree_set_page(zkm_pagetree_t* t, uint32_t page_index, const uint8_t* data, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(t && data, "null argument");
    ZKM_CHECK(page_index >= (zkm::PT_MAX_MEMORY >> 12) && page_index <= zkm::PT_ROOT_PAGE, "not a hash page index");
    zkm::PageTreeDev::Page p;
    memcpy(p.data(), data, zkm::PAGE_BYTES);
    t->t.hash_pages[page_index] = p;
    ZKM_API_END
}

// ---- the whole split_segment (emulator/src/state.rs:1477-1530) but the step loop: hashing on the device, the pre_* bookkeeping of
// InstrumentedState (:556-596) and the segment file.
struct zkm_splitter {
    zkm_pagetree tree;
    uint32_t pre_segment_id = 0, pre_pc = 0;
    uint8_t pre_image_id[32] = {0}, pre_hash_root[32] = {0};
    std::vector<std::vector<uint8_t>> pre_input;
    uint64_t pre_input_ptr = 0;
    std::vector<uint8_t> pre_public_values;
    uint64_t pre_public_values_ptr = 0;
};

int zkm_b200_splitter_create(zkm_splitter_t** out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(out, "null argument");
    *out = new zkm_splitter;
    ZKM_API_END
}
void zkm_b200_splitter_destroy(zkm_splitter_t* s) { delete s; }
zkm_pagetree_t* zkm_b200_splitter_pagetree(zkm_splitter_t* s) { return s ? &s->tree : nullptr; }
uint32_t zkm_b200_splitter_segment_count(const zkm_splitter_t* s) { return s ? s->pre_segment_id : 0; }

int zkm_b200_splitter_split(zkm_splitter_t* s, const zkm_split_state_t* st, int proof, char** segment_json_out, size_t* segment_json_len,
                            uint8_t* image_id_out, uint8_t* page_hash_root_out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(s && st && image_id_out && page_hash_root_out, "null argument");
    ZKM_CHECK(!proof || segment_json_out, "null argument");
    ZKM_CHECK((st->input_stream && st->input_stream_lens) || st->n_input_streams == 0, "null input stream");
    ZKM_CHECK(st->public_values_stream || st->public_values_stream_len == 0, "null public values stream");
    uint8_t image_id[32], root[32];
    // update_page_hash over wtrace[0], then compute_image_id(pc, registers)
    s->tree.t.split(st->dirty_page_indices, st->dirty_pages, st->n_dirty_pages, st->registers, st->pc, image_id, root, zkm::ctx().stream);
    if (segment_json_out) *segment_json_out = nullptr;
    if (proof) {
        // Segment { mem_image: get_input_image(), segment_id: pre_segment_id, pc: pre_pc, pre_hash_root, pre_image_id, image_id,
        //           end_pc: pc, step, page_hash_root, input_stream: pre_input, .. } written as serde_json (:1498-1519)
        std::vector<const uint8_t*> ptrs;
        std::vector<size_t> lens;
        for (auto& v : s->pre_input) { ptrs.push_back(v.data()); lens.push_back(v.size()); }
        zkm_segment_t seg = {};
        seg.page_indices = st->read_page_indices; seg.pages = st->read_pages; seg.n_pages = st->n_read_pages;
        seg.pc = s->pre_pc; seg.segment_id = s->pre_segment_id;
        memcpy(seg.pre_image_id, s->pre_image_id, 32); memcpy(seg.pre_hash_root, s->pre_hash_root, 32);
        memcpy(seg.image_id, image_id, 32); memcpy(seg.page_hash_root, root, 32);
        seg.end_pc = st->pc; seg.step = st->step;
        seg.input_stream = ptrs.data(); seg.input_stream_lens = lens.data(); seg.n_input_streams = ptrs.size();
        seg.input_stream_ptr = s->pre_input_ptr;
        seg.public_values_stream = s->pre_public_values.data(); seg.public_values_stream_len = s->pre_public_values.size();
        seg.public_values_stream_ptr = s->pre_public_values_ptr;
        char* jerr = nullptr;
        if (zkm_b200_segment_json(&seg, segment_json_out, segment_json_len, &jerr) != 0) {
            std::string m = jerr ? jerr : "segment json failed";
            if (jerr) zkm_b200_free_string(jerr);
            throw std::runtime_error(m);
        }
        s->pre_segment_id += 1;
    }
    s->pre_input.assign(st->n_input_streams, {});
    for (size_t k = 0; k < st->n_input_streams; k++) s->pre_input[k].assign(st->input_stream[k], st->input_stream[k] + st->input_stream_lens[k]);
    s->pre_input_ptr = st->input_stream_ptr;
    s->pre_public_values.assign(st->public_values_stream, st->public_values_stream + st->public_values_stream_len);
    s->pre_public_values_ptr = st->public_values_stream_ptr;
    s->pre_pc = st->pc;
    memcpy(s->pre_image_id, image_id, 32); memcpy(s->pre_hash_root, root, 32);
    memcpy(image_id_out, image_id, 32); memcpy(page_hash_root_out, root, 32);
    ZKM_API_END
}

}  // extern "C"
