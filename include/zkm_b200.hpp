// zkm_b200.hpp -- the host side of the drop-in in C++, above the C ABI of zkm_b200.h.
//
// The reference's host language (Rust) has no toolchain in this image, so next to the Rust source in shim/ (never compiled here)
// this header gives the same interface as compiled, tested code: the names, argument meaning and error behaviour of the
// reference's prover API for the path, with the reference's types rebuilt from the library's flat buffers.
//
//   reference (prover/src)                                         here (namespace zkm_b200)
//   config.rs:4-29        StarkConfig::standard_fast_config()       StarkConfig::standard_fast_config()
//   proof.rs:25-29        AllProof { stark_proofs: [..; 12], ctl_challenges, public_values }            AllProof
//   proof.rs:52-66        PublicValues { roots_before: MemRoots, roots_after: MemRoots, userdata }      PublicValues, MemRoots
//   proof.rs:178-201      StarkProof, StarkProofWithMetadata { init_challenger_state, proof }           same names
//   proof.rs:283-296      StarkOpeningSet { local_values, next_values, auxiliary_polys,
//                                           auxiliary_polys_next, ctl_zs_first, quotient_polys }        StarkOpeningSet
//   plonky2 fri/proof.rs  FriProof, FriQueryRound, FriInitialTreeProof, FriQueryStep                    same names
//   plonky2 hash/*        HashOut, MerkleCap, MerkleProof                                               same names
//   cross_table_lookup.rs:486-491  GrandProductChallenge { beta, gamma }, GrandProductChallengeSet      same names
//   prover.rs:130-140     prove_with_traces(all_stark, config, trace_poly_values, public_values, timing) -> Result<AllProof>
//                                                                   prove_with_traces(config, trace_poly_values, public_values, &timing)
//   prover.rs:86,146..    TimingTree scopes                          TimingTree { scopes: (depth, milliseconds, name) }
//   utils.rs:156-161      serde_json::to_string(&proof)              to_json(all_proof, table), public_values_json(all_proof)
// Errors (anyhow::Error / panics of the reference, with the reference's message texts) are thrown as zkm_b200::Error.
// F = GoldilocksField as its canonical u64, D = 2 (the only instantiation in the tree: utils.rs:34-36).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zkm_b200.h"

namespace zkm_b200 {

constexpr size_t NUM_TABLES = 12;                          // all_stark.rs:112
enum class Table : uint32_t {                              // all_stark.rs:97-110
    Arithmetic = 0, Cpu, Poseidon, PoseidonSponge, Keccak, KeccakSponge, ShaExtend, ShaExtendSponge, ShaCompress, ShaCompressSponge, Logic, Memory
};

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

typedef uint64_t F;                                        // GoldilocksField, canonical
struct Ext { F a = 0, b = 0; bool operator==(const Ext& o) const { return a == o.a && b == o.b; } };      // QuadraticExtension<F>
struct HashOut { std::array<F, 4> elements{}; bool operator==(const HashOut& o) const { return elements == o.elements; } };
typedef std::vector<HashOut> MerkleCap;
struct MerkleProof { std::vector<HashOut> siblings; };
struct PolynomialValues { std::vector<F> values; };       // one trace column on H
struct PolynomialCoeffsExt { std::vector<Ext> coeffs; };

struct StarkOpeningSet {
    std::vector<Ext> local_values, next_values, auxiliary_polys, auxiliary_polys_next;
    std::vector<F> ctl_zs_first;
    std::vector<Ext> quotient_polys;
};
struct FriInitialTreeProof { std::vector<std::pair<std::vector<F>, MerkleProof>> evals_proofs; };
struct FriQueryStep { std::vector<Ext> evals; MerkleProof merkle_proof; };
struct FriQueryRound { FriInitialTreeProof initial_trees_proof; std::vector<FriQueryStep> steps; };
struct FriProof {
    std::vector<MerkleCap> commit_phase_merkle_caps;
    std::vector<FriQueryRound> query_round_proofs;
    PolynomialCoeffsExt final_poly;
    F pow_witness = 0;
};
struct StarkProof {
    MerkleCap trace_cap, auxiliary_polys_cap, quotient_polys_cap;
    StarkOpeningSet openings;
    FriProof opening_proof;
    // proof.rs:203-211
    size_t recover_degree_bits(const zkm_stark_config_t& c) const {
        if (opening_proof.query_round_proofs.empty() || opening_proof.query_round_proofs[0].initial_trees_proof.evals_proofs.empty())
            throw Error("proof has no query rounds");
        return opening_proof.query_round_proofs[0].initial_trees_proof.evals_proofs[0].second.siblings.size() + c.cap_height - c.rate_bits;
    }
};
struct StarkProofWithMetadata { std::array<F, 12> init_challenger_state{}; StarkProof proof; };
struct GrandProductChallenge { F beta = 0, gamma = 0; };
struct GrandProductChallengeSet { std::vector<GrandProductChallenge> challenges; };
struct MemRoots { std::array<uint32_t, 8> root{}; };
struct PublicValues { MemRoots roots_before, roots_after; std::vector<uint8_t> userdata; };
struct AllProof {
    std::array<StarkProofWithMetadata, NUM_TABLES> stark_proofs;
    GrandProductChallengeSet ctl_challenges;
    PublicValues public_values;
    std::array<size_t, NUM_TABLES> degree_bits(const zkm_stark_config_t& c) const {       // proof.rs:31-37
        std::array<size_t, NUM_TABLES> d{};
        for (size_t t = 0; t < NUM_TABLES; t++) d[t] = stark_proofs[t].proof.recover_degree_bits(c);
        return d;
    }
};

struct StarkConfig {
    zkm_stark_config_t c;
    static StarkConfig standard_fast_config() { StarkConfig s; zkm_b200_standard_fast_config(&s.c); return s; }
};

// Device-time scopes of the last proof keyed by the reference's TimingTree scope strings.
struct TimingTree {
    struct Scope { int depth; double milliseconds; std::string name; };
    std::vector<Scope> scopes;
};

namespace detail {
inline void check(int rc, char* err) {
    if (rc == 0) return;
    std::string m = err ? err : "zkm_b200: unknown error";
    if (err) zkm_b200_free_string(err);
    throw Error(m);
}
constexpr uint64_t PROOF_MAGIC = 0x464F4F52504D4B5AULL;       // "ZKMPROOF"
struct Reader {
    const uint64_t* p; size_t n, pos = 0;
    uint64_t u() { if (pos >= n) throw Error("proof buffer truncated"); return p[pos++]; }
    const uint64_t* words(size_t k) { if (k > n - pos) throw Error("proof buffer truncated"); const uint64_t* r = p + pos; pos += k; return r; }
    std::vector<F> fs() { size_t k = (size_t)u(); const uint64_t* w = words(k); return std::vector<F>(w, w + k); }
    std::vector<Ext> exts() {
        size_t k = (size_t)u();
        if (k > (n - pos) / 2) throw Error("proof buffer truncated");
        const uint64_t* w = words(2 * k);
        std::vector<Ext> v(k);
        for (size_t i = 0; i < k; i++) { v[i].a = w[2 * i]; v[i].b = w[2 * i + 1]; }
        return v;
    }
    std::vector<HashOut> hashes() {
        size_t k = (size_t)u();
        if (k > (n - pos) / 4) throw Error("proof buffer truncated");
        const uint64_t* w = words(4 * k);
        std::vector<HashOut> v(k);
        for (size_t i = 0; i < k; i++) for (int j = 0; j < 4; j++) v[i].elements[j] = w[4 * i + j];
        return v;
    }
};
struct Writer {
    std::vector<uint64_t> w;
    void u(uint64_t x) { w.push_back(x); }
    void fs(const std::vector<F>& v) { u(v.size()); w.insert(w.end(), v.begin(), v.end()); }
    void exts(const std::vector<Ext>& v) { u(v.size()); for (auto& e : v) { u(e.a); u(e.b); } }
    void hashes(const std::vector<HashOut>& v) { u(v.size()); for (auto& h : v) for (F x : h.elements) u(x); }
};
inline StarkProofWithMetadata decode_table(Reader& r) {
    StarkProofWithMetadata t;
    const uint64_t* st = r.words(12);
    std::copy(st, st + 12, t.init_challenger_state.begin());
    StarkProof& p = t.proof;
    p.trace_cap = r.hashes(); p.auxiliary_polys_cap = r.hashes(); p.quotient_polys_cap = r.hashes();
    p.openings.local_values = r.exts(); p.openings.next_values = r.exts();
    p.openings.auxiliary_polys = r.exts(); p.openings.auxiliary_polys_next = r.exts();
    p.openings.ctl_zs_first = r.fs(); p.openings.quotient_polys = r.exts();
    size_t ncaps = (size_t)r.u();
    for (size_t i = 0; i < ncaps; i++) p.opening_proof.commit_phase_merkle_caps.push_back(r.hashes());
    size_t nq = (size_t)r.u();
    for (size_t q = 0; q < nq; q++) {
        FriQueryRound round;
        size_t no = (size_t)r.u();
        for (size_t o = 0; o < no; o++) { std::vector<F> leaf = r.fs(); MerkleProof mp{r.hashes()}; round.initial_trees_proof.evals_proofs.emplace_back(std::move(leaf), std::move(mp)); }
        size_t ns = (size_t)r.u();
        for (size_t s = 0; s < ns; s++) { FriQueryStep step; step.evals = r.exts(); step.merkle_proof.siblings = r.hashes(); round.steps.push_back(std::move(step)); }
        p.opening_proof.query_round_proofs.push_back(std::move(round));
    }
    p.opening_proof.final_poly.coeffs = r.exts();
    p.opening_proof.pow_witness = r.u();
    return t;
}
}  // namespace detail

// Rebuilds AllProof from the library's flat buffer (layout: zkm_b200.h "Proof buffer layout").
inline AllProof decode_all_proof(const uint64_t* buf, size_t words) {
    detail::Reader r{buf, words};
    if (r.u() != detail::PROOF_MAGIC) throw Error("bad proof magic");
    if (r.u() != 1) throw Error("bad proof version");
    if (r.u() != NUM_TABLES) throw Error("proof is not an AllStark proof");
    AllProof ap;
    size_t nch = (size_t)r.u();
    for (size_t k = 0; k < nch; k++) { GrandProductChallenge c; c.beta = r.u(); c.gamma = r.u(); ap.ctl_challenges.challenges.push_back(c); }
    for (auto& x : ap.public_values.roots_before.root) x = (uint32_t)r.u();
    for (auto& x : ap.public_values.roots_after.root) x = (uint32_t)r.u();
    for (F b : r.fs()) ap.public_values.userdata.push_back((uint8_t)b);
    for (size_t t = 0; t < NUM_TABLES; t++) ap.stark_proofs[t] = detail::decode_table(r);
    if (r.pos != words) throw Error("trailing data after proof");
    return ap;
}
// The inverse: AllProof -> flat buffer (what the wire-format entry points take).
inline std::vector<uint64_t> encode_all_proof(const AllProof& ap) {
    detail::Writer w;
    w.u(detail::PROOF_MAGIC); w.u(1); w.u(NUM_TABLES);
    w.u(ap.ctl_challenges.challenges.size());
    for (auto& c : ap.ctl_challenges.challenges) { w.u(c.beta); w.u(c.gamma); }
    for (uint32_t x : ap.public_values.roots_before.root) w.u(x);
    for (uint32_t x : ap.public_values.roots_after.root) w.u(x);
    w.u(ap.public_values.userdata.size());
    for (uint8_t b : ap.public_values.userdata) w.u(b);
    for (auto& t : ap.stark_proofs) {
        for (F x : t.init_challenger_state) w.u(x);
        const StarkProof& p = t.proof;
        w.hashes(p.trace_cap); w.hashes(p.auxiliary_polys_cap); w.hashes(p.quotient_polys_cap);
        w.exts(p.openings.local_values); w.exts(p.openings.next_values); w.exts(p.openings.auxiliary_polys);
        w.exts(p.openings.auxiliary_polys_next); w.fs(p.openings.ctl_zs_first); w.exts(p.openings.quotient_polys);
        w.u(p.opening_proof.commit_phase_merkle_caps.size());
        for (auto& c : p.opening_proof.commit_phase_merkle_caps) w.hashes(c);
        w.u(p.opening_proof.query_round_proofs.size());
        for (auto& q : p.opening_proof.query_round_proofs) {
            w.u(q.initial_trees_proof.evals_proofs.size());
            for (auto& ep : q.initial_trees_proof.evals_proofs) { w.fs(ep.first); w.hashes(ep.second.siblings); }
            w.u(q.steps.size());
            for (auto& s : q.steps) { w.exts(s.evals); w.hashes(s.merkle_proof.siblings); }
        }
        w.exts(p.opening_proof.final_poly.coeffs);
        w.u(p.opening_proof.pow_witness);
    }
    return std::move(w.w);
}

// zkm_b200_init: creates the device context; throws when no Blackwell device is present (there is no CPU fallback).
inline void init(int device = 0) { char* err = nullptr; detail::check(zkm_b200_init(device, &err), err); }

// prove_with_traces (prover.rs:130-140).  trace_poly_values[t] = the columns of table t (`Table` order), each 2^k values.
// timing: when given, receives the device-time scopes keyed by the reference's TimingTree scope strings.
inline AllProof prove_with_traces(const StarkConfig& config, const std::array<std::vector<PolynomialValues>, NUM_TABLES>& trace_poly_values,
                                  const PublicValues& public_values, TimingTree* timing = nullptr) {
    std::array<std::vector<const uint64_t*>, NUM_TABLES> ptrs;
    zkm_table_t tables[NUM_TABLES];
    for (size_t t = 0; t < NUM_TABLES; t++) {
        const auto& cols = trace_poly_values[t];
        if (cols.empty()) throw Error("null/empty table");
        const size_t n = cols[0].values.size();
        if (n == 0 || (n & (n - 1))) throw Error("trace length is not a power of two");
        uint32_t log_n = 0;
        while (((size_t)1 << log_n) < n) log_n++;
        for (auto& c : cols) {
            if (c.values.size() != n) throw Error("ragged table: columns of different lengths");
            ptrs[t].push_back(c.values.data());
        }
        tables[t] = zkm_table_t{ptrs[t].data(), (uint32_t)cols.size(), log_n};
    }
    if (timing) zkm_b200_timing_enable(1);
    uint64_t* out = nullptr; size_t words = 0; char* err = nullptr;
    int rc = zkm_b200_prove_with_traces(tables, public_values.roots_before.root.data(), public_values.roots_after.root.data(),
                                        public_values.userdata.data(), (uint32_t)public_values.userdata.size(), &config.c, &out, &words, &err);
    if (timing) {
        timing->scopes.clear();
        if (char* txt = zkm_b200_last_timing()) {
            for (char* line = txt; *line;) {
                char* end = strchr(line, '\n');
                std::string l(line, end ? (size_t)(end - line) : strlen(line));
                size_t t1 = l.find('\t'), t2 = l.find('\t', t1 + 1);
                if (t1 != std::string::npos && t2 != std::string::npos)
                    timing->scopes.push_back({std::stoi(l.substr(0, t1)), std::stod(l.substr(t1 + 1, t2 - t1 - 1)), l.substr(t2 + 1)});
                if (!end) break;
                line = end + 1;
            }
            zkm_b200_free_string(txt);
        }
        zkm_b200_timing_enable(0);
    }
    detail::check(rc, err);
    struct Free { uint64_t* p; ~Free() { zkm_b200_free(p); } } guard{out};
    AllProof ap = decode_all_proof(out, words);
    if (ap.public_values.userdata != public_values.userdata) throw Error("public values were not echoed back");
    return ap;
}

// serde_json::to_string(&all_proof.stark_proofs[table].proof) / of the public values (utils.rs:156-161, recursion/src/lib.rs:142-146).
inline std::string to_json(const AllProof& ap, Table table) {
    std::vector<uint64_t> buf = encode_all_proof(ap);
    char* json = nullptr; size_t len = 0; char* err = nullptr;
    detail::check(zkm_b200_proof_table_json(buf.data(), buf.size(), (uint32_t)table, &json, &len, &err), err);
    std::string s(json, len);
    zkm_b200_free_string(json);
    return s;
}
inline std::string public_values_json(const AllProof& ap) {
    std::vector<uint64_t> buf = encode_all_proof(ap);
    char* json = nullptr; size_t len = 0; char* err = nullptr;
    detail::check(zkm_b200_public_values_json(buf.data(), buf.size(), &json, &len, &err), err);
    std::string s(json, len);
    zkm_b200_free_string(json);
    return s;
}

// ---- proving straight from the witness (`Traces`, witness/traces.rs:46-60): every table but Cpu crosses the ABI as its operation
// log and is generated on the device (zkm_b200_prove_with_ops), the Cpu rows go row-major.  The C++ counterpart of shim/src/b200_ops.rs.
struct MemoryAddress { uint64_t context = 0, segment = 0, virt = 0; };                     // witness/memory.rs:27-31
enum class MemoryOpKind { Read, Write };
struct MemoryOp { bool filter = true; uint64_t timestamp = 0; MemoryAddress address; MemoryOpKind kind = MemoryOpKind::Read; uint32_t value = 0; };
struct ArithmeticOperation { uint32_t row_filter = 0, input0 = 0, input1 = 0; };            // BinaryOperator::row_filter() = the IS_* column
enum class LogicOp : uint32_t { And = 0, Or, Xor, Nor };                                    // logic.rs:84-89
struct LogicOperation { LogicOp operator_ = LogicOp::And; uint32_t input0 = 0, input1 = 0; };
struct SpongeOp { std::vector<MemoryAddress> base_address; uint64_t timestamp = 0; std::vector<uint8_t> input; };   // Keccak / PoseidonSpongeOp
struct ShaExtendSpongeOp { std::vector<MemoryAddress> base_address; uint64_t timestamp = 0; std::array<uint8_t, 16> input{}; uint32_t i = 0;
                           MemoryAddress output_address; };
struct ShaCompressInput { std::array<uint8_t, 41> input{}; MemoryAddress w_i_address; uint64_t timestamp = 0; };    // one ROW of ShaCompress
struct ShaCompressSpongeOp { std::vector<MemoryAddress> base_address; uint64_t timestamp = 0; std::array<uint8_t, 32> input{};
                             std::vector<std::array<uint8_t, 4>> w_i_s; };
struct Traces {
    std::vector<ArithmeticOperation> arithmetic_ops;
    std::vector<F> cpu;                                                                       // rows x 259, row-major, padded to a power of two
    std::vector<LogicOperation> logic_ops;
    std::vector<MemoryOp> memory_ops;
    std::vector<std::pair<std::array<F, 12>, uint64_t>> poseidon_inputs;
    std::vector<SpongeOp> poseidon_sponge_ops;
    std::vector<std::pair<std::array<uint64_t, 25>, uint64_t>> keccak_inputs;
    std::vector<SpongeOp> keccak_sponge_ops;
    std::vector<std::pair<std::array<uint8_t, 16>, uint64_t>> sha_extend_inputs;
    std::vector<ShaExtendSpongeOp> sha_extend_sponge_ops;
    std::vector<ShaCompressInput> sha_compress_inputs;
    std::vector<ShaCompressSpongeOp> sha_compress_sponge_ops;
};
struct OpLog { std::vector<uint64_t> words; size_t n_ops = 0; };

namespace detail {
inline uint64_t le_u32(const uint8_t* b) { return (uint64_t)b[0] | (uint64_t)b[1] << 8 | (uint64_t)b[2] << 16 | (uint64_t)b[3] << 24; }
inline OpLog sponge_log(const std::vector<SpongeOp>& ops) {
    OpLog log;
    log.words.push_back(0);
    for (const SpongeOp& op : ops) {
        if (op.base_address.empty()) throw Error("sponge operation without addresses");
        for (uint64_t w : {op.base_address[0].context, op.base_address[0].segment, op.timestamp, (uint64_t)op.input.size(), (uint64_t)op.base_address.size()})
            log.words.push_back(w);
        for (const MemoryAddress& a : op.base_address) log.words.push_back(a.virt);
        for (size_t i = 0; i < op.input.size(); i += 8) {
            uint64_t w = 0;
            for (size_t j = 0; j < 8 && i + j < op.input.size(); j++) w |= (uint64_t)op.input[i + j] << (8 * j);
            log.words.push_back(w);
        }
        log.n_ops++;
    }
    log.words[0] = log.words.size();
    return log;
}
}  // namespace detail

// The eleven operation logs, indexed by `Table` (formats: zkm_b200.h next to zkm_op_log_t); logs[Cpu] stays empty.
inline std::array<OpLog, NUM_TABLES> op_logs(const Traces& t) {
    std::array<OpLog, NUM_TABLES> logs;
    auto at = [&](Table x) -> OpLog& { return logs[(size_t)x]; };
    for (auto& o : t.arithmetic_ops) { auto& l = at(Table::Arithmetic); l.words.insert(l.words.end(), {o.row_filter, o.input0, o.input1}); l.n_ops++; }
    for (auto& o : t.logic_ops) { auto& l = at(Table::Logic); l.words.insert(l.words.end(), {(uint64_t)o.operator_, o.input0, o.input1}); l.n_ops++; }
    for (auto& m : t.memory_ops) {
        auto& l = at(Table::Memory);
        l.words.insert(l.words.end(), {m.address.context, m.address.segment, m.address.virt, m.timestamp, (uint64_t)(m.kind == MemoryOpKind::Read), m.value,
                                       (uint64_t)m.filter});
        l.n_ops++;
    }
    for (auto& p : t.poseidon_inputs) { auto& l = at(Table::Poseidon); l.words.insert(l.words.end(), p.first.begin(), p.first.end()); l.words.push_back(p.second); l.n_ops++; }
    for (auto& k : t.keccak_inputs) { auto& l = at(Table::Keccak); l.words.insert(l.words.end(), k.first.begin(), k.first.end()); l.words.push_back(k.second); l.n_ops++; }
    at(Table::PoseidonSponge) = detail::sponge_log(t.poseidon_sponge_ops);
    at(Table::KeccakSponge) = detail::sponge_log(t.keccak_sponge_ops);
    for (auto& e : t.sha_extend_inputs) {
        auto& l = at(Table::ShaExtend);
        for (int k = 0; k < 4; k++) l.words.push_back(detail::le_u32(e.first.data() + 4 * k));
        l.words.push_back(e.second); l.n_ops++;
    }
    for (auto& o : t.sha_extend_sponge_ops) {
        auto& l = at(Table::ShaExtendSponge);
        if (o.base_address.size() < 4) throw Error("sha extend sponge operation needs four addresses");
        l.words.push_back(o.i);
        for (int k = 0; k < 4; k++) l.words.push_back(detail::le_u32(o.input.data() + 4 * k));
        for (int k = 0; k < 4; k++) l.words.push_back(o.base_address[k].virt);
        l.words.insert(l.words.end(), {o.output_address.virt, o.base_address[0].context, o.base_address[0].segment, o.timestamp});
        l.n_ops++;
    }
    for (auto& r : t.sha_compress_inputs) {
        auto& l = at(Table::ShaCompress);
        for (int k = 0; k < 10; k++) l.words.push_back(detail::le_u32(r.input.data() + 4 * k));
        l.words.insert(l.words.end(), {(uint64_t)r.input[40], r.w_i_address.virt, r.w_i_address.segment, r.w_i_address.context, r.timestamp});
        l.n_ops++;
    }
    for (auto& o : t.sha_compress_sponge_ops) {
        auto& l = at(Table::ShaCompressSponge);
        if (o.base_address.size() < 9 || o.w_i_s.size() != 64) throw Error("sha compress sponge operation needs nine addresses and 64 message words");
        for (int k = 0; k < 8; k++) l.words.push_back(detail::le_u32(o.input.data() + 4 * k));
        for (auto& w : o.w_i_s) l.words.push_back(detail::le_u32(w.data()));
        for (int k = 0; k < 8; k++) l.words.push_back(o.base_address[k].virt);
        const MemoryAddress& ws = o.base_address[8];
        l.words.insert(l.words.end(), {ws.virt, ws.segment, ws.context, o.base_address[0].context, o.base_address[0].segment, o.timestamp});
        l.n_ops++;
    }
    return logs;
}

// Traces::into_tables + prove_with_traces in one device-side call (prover.rs:58-128 call chain).
inline AllProof prove_from_traces(const StarkConfig& config, const Traces& traces, const PublicValues& public_values) {
    constexpr uint32_t CPU_COLUMNS = 259;
    const size_t rows = traces.cpu.size() / CPU_COLUMNS;
    if (rows * CPU_COLUMNS != traces.cpu.size() || rows < 64 || (rows & (rows - 1))) throw Error("the Cpu rows must be padded to a power of two");
    uint32_t log_n = 0;
    while (((size_t)1 << log_n) < rows) log_n++;
    std::array<OpLog, NUM_TABLES> logs = op_logs(traces);
    zkm_table_t tables[NUM_TABLES] = {};
    zkm_table_rows_t row_tables[NUM_TABLES] = {};
    zkm_op_log_t c_logs[NUM_TABLES] = {};
    row_tables[(size_t)Table::Cpu] = zkm_table_rows_t{traces.cpu.data(), CPU_COLUMNS, log_n};
    static const uint64_t empty_log = 0;
    for (size_t t = 0; t < NUM_TABLES; t++) {
        if (t == (size_t)Table::Cpu) continue;
        c_logs[t].ops = logs[t].words.empty() ? &empty_log : logs[t].words.data();        // a non-null pointer selects the device generator
        c_logs[t].n_ops = logs[t].n_ops;
    }
    uint64_t* out = nullptr; size_t words = 0; char* err = nullptr;
    detail::check(zkm_b200_prove_with_ops(tables, row_tables, c_logs, public_values.roots_before.root.data(), public_values.roots_after.root.data(),
                                          public_values.userdata.data(), (uint32_t)public_values.userdata.size(), &config.c, &out, &words, &err), err);
    struct Free { uint64_t* p; ~Free() { zkm_b200_free(p); } } guard{out};
    return decode_all_proof(out, words);
}

}  // namespace zkm_b200
