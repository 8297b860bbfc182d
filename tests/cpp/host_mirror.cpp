// Test driver for include/zkm_b200.hpp (the C++ host-side mirror of the reference's prover API).  Built and run by
// tests/test_cpp_host.py.
//   host_mirror decode <proof.bin> <table>          CPU: decode -> re-encode round trip, shape summary, JSON of one table
//   host_mirror prove <out.bin> <log heights x 12>  GPU: prove_with_traces over the synthetic traces, proof written to <out.bin>
//   host_mirror errors                              CPU: the error paths that need no device
//   host_mirror oplogs <traces.bin> <logs.bin>      CPU: a typed `Traces` (dumped by the test, one field per word) -> op_logs() ->
//                                                   per table: n_ops, words, the words
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include "zkm_b200.hpp"

using namespace zkm_b200;

static std::vector<uint64_t> read_words(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(std::string("cannot open ") + path);
    std::vector<char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<uint64_t> w(bytes.size() / 8);
    memcpy(w.data(), bytes.data(), w.size() * 8);
    return w;
}

// The typed Traces dump of tests/test_cpp_host.py: every scalar one word, every byte one word, sections in `Traces` field order.
struct Words {
    std::vector<uint64_t> w; size_t at = 0;
    uint64_t next() { if (at >= w.size()) throw Error("traces dump truncated"); return w[at++]; }
    MemoryAddress addr() { MemoryAddress a; a.context = next(); a.segment = next(); a.virt = next(); return a; }
    std::vector<MemoryAddress> addrs() { std::vector<MemoryAddress> v(next()); for (auto& a : v) a = addr(); return v; }
    template <size_t N> std::array<uint8_t, N> bytes() { std::array<uint8_t, N> b; for (auto& x : b) x = (uint8_t)next(); return b; }
};
static Traces read_traces(const char* path) {
    Words in{read_words(path)};
    Traces t;
    t.arithmetic_ops.resize(in.next());
    for (auto& o : t.arithmetic_ops) { o.row_filter = (uint32_t)in.next(); o.input0 = (uint32_t)in.next(); o.input1 = (uint32_t)in.next(); }
    t.logic_ops.resize(in.next());
    for (auto& o : t.logic_ops) { o.operator_ = (LogicOp)in.next(); o.input0 = (uint32_t)in.next(); o.input1 = (uint32_t)in.next(); }
    t.memory_ops.resize(in.next());
    for (auto& m : t.memory_ops) {
        m.address = in.addr(); m.timestamp = in.next(); m.kind = in.next() ? MemoryOpKind::Read : MemoryOpKind::Write;
        m.value = (uint32_t)in.next(); m.filter = in.next() != 0;
    }
    t.poseidon_inputs.resize(in.next());
    for (auto& p : t.poseidon_inputs) { for (auto& x : p.first) x = in.next(); p.second = in.next(); }
    for (auto* ops : {&t.poseidon_sponge_ops, &t.keccak_sponge_ops}) {
        ops->resize(in.next());
        for (auto& o : *ops) { o.base_address = in.addrs(); o.timestamp = in.next(); o.input.resize(in.next()); for (auto& b : o.input) b = (uint8_t)in.next(); }
    }
    t.keccak_inputs.resize(in.next());
    for (auto& k : t.keccak_inputs) { for (auto& x : k.first) x = in.next(); k.second = in.next(); }
    t.sha_extend_inputs.resize(in.next());
    for (auto& e : t.sha_extend_inputs) { e.first = in.bytes<16>(); e.second = in.next(); }
    t.sha_extend_sponge_ops.resize(in.next());
    for (auto& o : t.sha_extend_sponge_ops) { o.base_address = in.addrs(); o.timestamp = in.next(); o.input = in.bytes<16>(); o.i = (uint32_t)in.next(); o.output_address = in.addr(); }
    t.sha_compress_inputs.resize(in.next());
    for (auto& r : t.sha_compress_inputs) { r.input = in.bytes<41>(); r.w_i_address = in.addr(); r.timestamp = in.next(); }
    t.sha_compress_sponge_ops.resize(in.next());
    for (auto& o : t.sha_compress_sponge_ops) {
        o.base_address = in.addrs(); o.timestamp = in.next(); o.input = in.bytes<32>();
        o.w_i_s.resize(in.next());
        for (auto& w : o.w_i_s) w = in.bytes<4>();
    }
    if (in.at != in.w.size()) throw Error("traces dump has trailing words");
    return t;
}

int main(int argc, char** argv) {
    try {
        const std::string mode = argc > 1 ? argv[1] : "";
        if (mode == "oplogs" && argc == 4) {
            Traces t = read_traces(argv[2]);
            std::array<OpLog, NUM_TABLES> logs = op_logs(t);
            std::ofstream out(argv[3], std::ios::binary);
            for (const OpLog& l : logs) {
                const uint64_t head[2] = {l.n_ops, l.words.size()};
                out.write((const char*)head, 16);
                out.write((const char*)l.words.data(), (std::streamsize)(l.words.size() * 8));
            }
            // the shape errors of the typed operations are reported, not undefined behaviour
            Traces bad = t;
            bad.sha_compress_sponge_ops.emplace_back();
            try { op_logs(bad); std::puts("SHORT OP NOT DETECTED"); return 1; } catch (const Error& e) { std::printf("short: %s\n", e.what()); }
            bad = Traces{};
            bad.cpu.assign(259 * 48, 0);
            try { prove_from_traces(StarkConfig::standard_fast_config(), bad, PublicValues{}); return 1; } catch (const Error& e) { std::printf("cpu: %s\n", e.what()); }
            return 0;
        }
        if (mode == "decode" && argc == 4) {
            std::vector<uint64_t> buf = read_words(argv[2]);
            AllProof ap = decode_all_proof(buf.data(), buf.size());
            if (encode_all_proof(ap) != buf) { std::puts("ROUNDTRIP MISMATCH"); return 1; }
            StarkConfig cfg = StarkConfig::standard_fast_config();
            std::printf("challenges %zu userdata %zu\n", ap.ctl_challenges.challenges.size(), ap.public_values.userdata.size());
            auto db = ap.degree_bits(cfg.c);
            for (size_t t = 0; t < NUM_TABLES; t++) {
                const StarkProof& p = ap.stark_proofs[t].proof;
                std::printf("table %zu degree_bits %zu caps %zu %zu %zu local %zu aux %zu zs %zu quot %zu queries %zu steps %zu final %zu pow %llu\n", t, db[t],
                            p.trace_cap.size(), p.auxiliary_polys_cap.size(), p.quotient_polys_cap.size(), p.openings.local_values.size(),
                            p.openings.auxiliary_polys.size(), p.openings.ctl_zs_first.size(), p.openings.quotient_polys.size(),
                            p.opening_proof.query_round_proofs.size(), p.opening_proof.query_round_proofs[0].steps.size(),
                            p.opening_proof.final_poly.coeffs.size(), (unsigned long long)p.opening_proof.pow_witness);
            }
            std::printf("JSON %s\n", to_json(ap, (Table)std::atoi(argv[3])).c_str());
            std::printf("PV %s\n", public_values_json(ap).c_str());
            // a truncated buffer is an error, not a crash
            try { decode_all_proof(buf.data(), buf.size() - 5); std::puts("TRUNCATION NOT DETECTED"); return 1; } catch (const Error& e) { std::printf("truncated: %s\n", e.what()); }
            // lengths come from the buffer: counts whose byte size wraps must be errors as well (any word of the first table)
            size_t refused = 0;
            for (size_t at = 2; at < std::min<size_t>(buf.size(), 1500); at++)
                for (uint64_t k : {~0ULL, 1ULL << 63, 1ULL << 62, (1ULL << 61) + 3}) {
                    std::vector<uint64_t> bad = buf;
                    bad[at] = k;
                    try { decode_all_proof(bad.data(), bad.size()); } catch (const std::exception&) { refused++; }
                }
            std::printf("hostile lengths refused: %zu\n", refused);
            return 0;
        }
        if (mode == "prove" && argc == 3 + (int)NUM_TABLES) {
            init(0);
            StarkConfig cfg = StarkConfig::standard_fast_config();
            uint32_t nt = 0, ncols[NUM_TABLES];
            char* err = nullptr;
            detail::check(zkm_b200_system_shape(0, &nt, ncols, NUM_TABLES, &err), err);
            std::array<std::vector<PolynomialValues>, NUM_TABLES> traces;
            for (size_t t = 0; t < NUM_TABLES; t++) {
                const uint32_t log_n = (uint32_t)std::atoi(argv[3 + t]);
                const size_t n = (size_t)1 << log_n;
                std::vector<uint64_t> flat((size_t)ncols[t] * n);
                detail::check(zkm_b200_synth_trace(0, (uint32_t)t, log_n, 0x5EED000000000000ULL | ((uint64_t)t << 16), flat.data(), &err), err);
                traces[t].resize(ncols[t]);
                for (uint32_t c = 0; c < ncols[t]; c++) traces[t][c].values.assign(flat.begin() + (size_t)c * n, flat.begin() + (size_t)(c + 1) * n);
            }
            PublicValues pv;
            for (int i = 0; i < 8; i++) { pv.roots_before.root[i] = 1 + i; pv.roots_after.root[i] = 11 + i; }
            pv.userdata.assign(32, 0);
            TimingTree timing;
            AllProof ap = prove_with_traces(cfg, traces, pv, &timing);
            std::vector<uint64_t> buf = encode_all_proof(ap);
            std::ofstream(argv[2], std::ios::binary).write((const char*)buf.data(), (std::streamsize)(buf.size() * 8));
            std::printf("proved %zu words, %zu timing scopes, first scope: %s\n", buf.size(), timing.scopes.size(),
                        timing.scopes.empty() ? "-" : timing.scopes[0].name.c_str());
            // a column of another length is refused before anything is uploaded
            traces[(size_t)Table::Logic][1].values.resize(3);
            try { prove_with_traces(cfg, traces, pv); std::puts("RAGGED NOT DETECTED"); return 1; } catch (const Error& e) { std::printf("ragged: %s\n", e.what()); }
            return 0;
        }
        if (mode == "errors") {
            StarkConfig cfg = StarkConfig::standard_fast_config();
            std::printf("config %u %u %u %u %u %u %u\n", cfg.c.rate_bits, cfg.c.cap_height, cfg.c.pow_bits, cfg.c.num_queries, cfg.c.num_challenges,
                        cfg.c.arity_bits, cfg.c.final_poly_bits);
            uint64_t junk[4] = {1, 2, 3, 4};
            try { decode_all_proof(junk, 4); return 1; } catch (const Error& e) { std::printf("junk: %s\n", e.what()); }
            std::array<std::vector<PolynomialValues>, NUM_TABLES> traces;
            try { prove_with_traces(cfg, traces, PublicValues{}); return 1; } catch (const Error& e) { std::printf("empty: %s\n", e.what()); }
            return 0;
        }
        std::fprintf(stderr, "usage: host_mirror decode <proof.bin> <table> | prove <out.bin> <12 log heights> | errors\n");
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
