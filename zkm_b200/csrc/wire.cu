// Proof wire format: the serde_json text of the reference's proof structs, produced from the flat proof buffer.
// Replaces `serde_json::to_string(&stark_proof)` for `StarkProof<F, C, D>` (reference prover/src/proof.rs:177-189,
// #[derive(Serialize)]; used for the proof-size log at prover/examples/utils/src/utils.rs:156-161) and
// `serde_json::to_string(&public_values)` for `PublicValues` (proof.rs:52-66; written to disk at recursion/src/lib.rs:142-146).
// serde's derived encodings of the plonky2 0.1.4 types (not in /root/reference; restated):
//   GoldilocksField(u64)             newtype -> the number (canonical u64, decimal)
//   QuadraticExtension([F; 2])       newtype -> [a, b]
//   HashOut { elements: [F; 4] }     -> {"elements":[e0,e1,e2,e3]}
//   MerkleCap(Vec<HashOut>)          newtype -> [hash, ...]
//   MerkleProof { siblings }         -> {"siblings":[hash, ...]}
//   PolynomialCoeffs { coeffs }      -> {"coeffs":[...]}
//   FriInitialTreeProof { evals_proofs: Vec<(Vec<F>, MerkleProof)> }  -> {"evals_proofs":[[[...],{"siblings":[...]}], ...]}
//   FriQueryStep { evals, merkle_proof }, FriQueryRound { initial_trees_proof, steps },
//   FriProof { commit_phase_merkle_caps, query_round_proofs, final_poly, pow_witness }
// serde_json::to_string is compact (no whitespace), fields in declaration order.  Host-only code: needs no device.
#include "../../include/zkm_b200.h"
#include "dev.cuh"
#include <cstring>
#include <cstdlib>
#include <string>

namespace zkm {
namespace {

struct Rd {
    const u64* p; size_t n, pos = 0;
    u64 u() { if (pos >= n) throw std::runtime_error("proof buffer truncated"); return p[pos++]; }
    // `k` items of `unit` words; the counts come from the buffer, so neither pos + k nor k * unit may be trusted not to wrap
    const u64* words(size_t k, size_t unit = 1) {
        if (k > (n - pos) / unit) throw std::runtime_error("proof buffer truncated");
        const u64* r = p + pos; pos += k * unit; return r;
    }
};
struct Js {
    std::string s;
    void num(u64 x) { char b[24]; int k = snprintf(b, sizeof b, "%llu", (unsigned long long)x); s.append(b, k); }
    void ext(const u64* w) { s += '['; num(w[0]); s += ','; num(w[1]); s += ']'; }
    void hash(const u64* w) { s += "{\"elements\":["; for (int i = 0; i < 4; i++) { if (i) s += ','; num(w[i]); } s += "]}"; }
    template <class F> void list(size_t n, F item) { s += '['; for (size_t i = 0; i < n; i++) { if (i) s += ','; item(i); } s += ']'; }
    void fs(const u64* w, size_t n) { list(n, [&](size_t i) { num(w[i]); }); }
    void exts(const u64* w, size_t n) { list(n, [&](size_t i) { ext(w + 2 * i); }); }
    void hashes(const u64* w, size_t n) { list(n, [&](size_t i) { hash(w + 4 * i); }); }
    void key(const char* k) { s += '"'; s += k; s += "\":"; }
};

// positions the reader at the StarkProofWithMetadata of `table` and returns the header fields
struct Header { size_t num_tables; std::vector<u64> pv; };
static const u64 MAGIC = 0x464F4F52504D4B5AULL;

void skip_table(Rd& r) {
    r.words(12);
    for (int c = 0; c < 3; c++) { size_t k = r.u(); r.words(k, 4); }
    for (int v = 0; v < 4; v++) { size_t k = r.u(); r.words(k, 2); }
    { size_t k = r.u(); r.words(k); }
    { size_t k = r.u(); r.words(k, 2); }
    size_t ncaps = r.u();
    for (size_t i = 0; i < ncaps; i++) { size_t k = r.u(); r.words(k, 4); }
    size_t nq = r.u();
    for (size_t q = 0; q < nq; q++) {
        size_t no = r.u();
        for (size_t o = 0; o < no; o++) { size_t k = r.u(); r.words(k); k = r.u(); r.words(k, 4); }
        size_t ns = r.u();
        for (size_t st = 0; st < ns; st++) { size_t k = r.u(); r.words(k, 2); k = r.u(); r.words(k, 4); }
    }
    { size_t k = r.u(); r.words(k, 2); }
    r.u();
}

void vec_field(Rd& r, Js& j, const char* name, int unit) {
    j.key(name);
    size_t k = r.u();
    const u64* w = r.words(k, (size_t)unit);
    if (unit == 1) j.fs(w, k); else if (unit == 2) j.exts(w, k); else j.hashes(w, k);
}

void stark_proof_json(Rd& r, Js& j) {
    r.words(12);                                            // init_challenger_state: metadata, not part of StarkProof
    j.s += '{';
    vec_field(r, j, "trace_cap", 4); j.s += ',';
    vec_field(r, j, "auxiliary_polys_cap", 4); j.s += ',';
    vec_field(r, j, "quotient_polys_cap", 4); j.s += ',';
    j.key("openings"); j.s += '{';
    vec_field(r, j, "local_values", 2); j.s += ',';
    vec_field(r, j, "next_values", 2); j.s += ',';
    vec_field(r, j, "auxiliary_polys", 2); j.s += ',';
    vec_field(r, j, "auxiliary_polys_next", 2); j.s += ',';
    vec_field(r, j, "ctl_zs_first", 1); j.s += ',';
    vec_field(r, j, "quotient_polys", 2);
    j.s += "},";
    j.key("opening_proof"); j.s += '{';
    j.key("commit_phase_merkle_caps");
    size_t ncaps = r.u();
    j.list(ncaps, [&](size_t) { size_t k = r.u(); j.hashes(r.words(k, 4), k); });
    j.s += ',';
    j.key("query_round_proofs");
    size_t nq = r.u();
    j.list(nq, [&](size_t) {
        j.s += '{';
        j.key("initial_trees_proof"); j.s += "{\"evals_proofs\":";
        size_t no = r.u();
        j.list(no, [&](size_t) {
            j.s += '[';
            size_t k = r.u(); j.fs(r.words(k), k);
            j.s += ",{\"siblings\":";
            k = r.u(); j.hashes(r.words(k, 4), k);
            j.s += "}]";
        });
        j.s += "},";
        j.key("steps");
        size_t ns = r.u();
        j.list(ns, [&](size_t) {
            j.s += "{\"evals\":";
            size_t k = r.u(); j.exts(r.words(k, 2), k);
            j.s += ",\"merkle_proof\":{\"siblings\":";
            k = r.u(); j.hashes(r.words(k, 4), k);
            j.s += "}}";
        });
        j.s += '}';
    });
    j.s += ',';
    j.key("final_poly"); j.s += "{\"coeffs\":";
    { size_t k = r.u(); j.exts(r.words(k, 2), k); }
    j.s += "},";
    j.key("pow_witness"); j.num(r.u());
    j.s += "}}";
}

char* dup(const std::string& s) {
    char* m = (char*)malloc(s.size() + 1);
    if (!m) throw std::runtime_error("out of host memory");
    memcpy(m, s.c_str(), s.size() + 1);
    return m;
}
int fail(char** err, const std::exception& e) {
    if (err) { const char* w = e.what(); size_t n = strlen(w); char* m = (char*)malloc(n + 1); if (m) memcpy(m, w, n + 1); *err = m; }
    return -1;
}
// reads the header up to the first table; returns num_tables and leaves the public values in the Js if asked
size_t read_header(Rd& r, Js* pv) {
    if (r.u() != MAGIC) throw std::runtime_error("bad proof magic");
    if (r.u() != 1) throw std::runtime_error("bad proof version");
    size_t nt = r.u();
    size_t nch = r.u();
    r.words(nch, 2);
    const u64* rb = r.words(8);
    const u64* ra = r.words(8);
    size_t nu = r.u();
    const u64* ud = r.words(nu);
    if (pv) {
        pv->s += "{\"roots_before\":{\"root\":"; pv->fs(rb, 8);
        pv->s += "},\"roots_after\":{\"root\":"; pv->fs(ra, 8);
        pv->s += "},\"userdata\":"; pv->fs(ud, nu); pv->s += '}';
    }
    return nt;
}

}  // namespace
}  // namespace zkm

using namespace zkm;
extern "C" {

int zkm_b200_proof_table_json(const uint64_t* proof, size_t proof_words, uint32_t table, char** json_out, size_t* json_len, char** err) {
    if (err) *err = nullptr;
    try {
        ZKM_CHECK(proof && json_out, "null argument");
        Rd r{proof, proof_words};
        size_t nt = read_header(r, nullptr);
        ZKM_CHECK(table < nt, "table index out of range");
        for (uint32_t t = 0; t < table; t++) skip_table(r);
        Js j;
        j.s.reserve(1 << 20);
        stark_proof_json(r, j);
        if (json_len) *json_len = j.s.size();
        *json_out = dup(j.s);
    } catch (const std::exception& e) { return fail(err, e); }
    return 0;
}

int zkm_b200_public_values_json(const uint64_t* proof, size_t proof_words, char** json_out, size_t* json_len, char** err) {
    if (err) *err = nullptr;
    try {
        ZKM_CHECK(proof && json_out, "null argument");
        Rd r{proof, proof_words};
        Js j;
        read_header(r, &j);
        if (json_len) *json_len = j.s.size();
        *json_out = dup(j.s);
    } catch (const std::exception& e) { return fail(err, e); }
    return 0;
}

// Segment (emulator/src/state.rs:33-48, #[derive(Serialize)]): the file split_segment writes per segment (:1498-1503) and
// `prove_segments` reads back (prover/examples/utils/src/utils.rs).  mem_image is get_input_image (memory.rs:524-538): every
// word of every page the segment read, keyed by address, value = u32::from_le_bytes of the page bytes; a BTreeMap<u32, u32>
// serialises as an object with decimal string keys in ascending key order.
int zkm_b200_segment_json(const zkm_segment_t* seg, char** json_out, size_t* json_len, char** err) {
    if (err) *err = nullptr;
    try {
        ZKM_CHECK(seg && json_out, "null argument");
        ZKM_CHECK((seg->page_indices && seg->pages) || seg->n_pages == 0, "null pages");
        Js j;
        j.s.reserve(64 + seg->n_pages * 4096 * 6);
        j.s += '{';
        j.key("mem_image"); j.s += '{';
        bool first = true;
        for (size_t k = 0; k < seg->n_pages; k++) {
            ZKM_CHECK(k == 0 || seg->page_indices[k] > seg->page_indices[k - 1], "page indices must be strictly ascending (BTreeMap order)");
            ZKM_CHECK(seg->page_indices[k] < (1u << 20), "page index out of range");
            const uint8_t* pg = seg->pages + k * 4096;
            for (uint32_t i = 0; i < 1024; i++) {
                uint32_t w; memcpy(&w, pg + 4 * i, 4);
                if (!first) j.s += ',';
                first = false;
                j.s += '"'; j.num(((u64)seg->page_indices[k] << 12) + 4 * i); j.s += "\":"; j.num(w);
            }
        }
        j.s += '}';
        auto bytes32 = [&](const char* name, const uint8_t* b) { j.s += ','; j.key(name); j.list(32, [&](size_t i) { j.num(b[i]); }); };
        j.s += ','; j.key("pc"); j.num(seg->pc);
        j.s += ','; j.key("segment_id"); j.num(seg->segment_id);
        bytes32("pre_image_id", seg->pre_image_id); bytes32("pre_hash_root", seg->pre_hash_root);
        bytes32("image_id", seg->image_id); bytes32("page_hash_root", seg->page_hash_root);
        j.s += ','; j.key("end_pc"); j.num(seg->end_pc);
        j.s += ','; j.key("step"); j.num(seg->step);
        j.s += ','; j.key("input_stream");
        ZKM_CHECK((seg->input_stream && seg->input_stream_lens) || seg->n_input_streams == 0, "null input stream");
        j.list(seg->n_input_streams, [&](size_t k) {
            const uint8_t* b = seg->input_stream[k];
            ZKM_CHECK(b || seg->input_stream_lens[k] == 0, "null input stream");
            j.list(seg->input_stream_lens[k], [&](size_t i) { j.num(b[i]); });
        });
        j.s += ','; j.key("input_stream_ptr"); j.num(seg->input_stream_ptr);
        j.s += ','; j.key("public_values_stream");
        ZKM_CHECK(seg->public_values_stream || seg->public_values_stream_len == 0, "null public values stream");
        j.list(seg->public_values_stream_len, [&](size_t i) { j.num(seg->public_values_stream[i]); });
        j.s += ','; j.key("public_values_stream_ptr"); j.num(seg->public_values_stream_ptr);
        j.s += '}';
        if (json_len) *json_len = j.s.size();
        *json_out = dup(j.s);
    } catch (const std::exception& e) { return fail(err, e); }
    return 0;
}

}  // extern "C"
