// Test driver for include/zkm_b200.hpp (the C++ host-side mirror of the reference's prover API).  Built and run by
// tests/test_cpp_host.py.
//   host_mirror decode <proof.bin> <table>          CPU: decode -> re-encode round trip, shape summary, JSON of one table
//   host_mirror prove <out.bin> <log heights x 12>  GPU: prove_with_traces over the synthetic traces, proof written to <out.bin>
//   host_mirror errors                              CPU: the error paths that need no device
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

int main(int argc, char** argv) {
    try {
        const std::string mode = argc > 1 ? argv[1] : "";
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
