/* zkm_b200 — C ABI of the B200-native STARK proving path for zkMIPS/zkm.
 *
 * The reference has no FFI for this path (it is a plain Rust call chain); this header is what a
 * `#[cfg(feature = "b200")]` shim inside reference prover/src/prover.rs would bind (INTEGRATION.md
 * shows the Rust `extern "C"` block).  Conventions follow the only FFI that exists in the reference
 * tree, the Groth16 wrapper (recursion/src/snark/snarks.rs:7-20,49-57): every entry point returns an
 * int status (0 = ok, -1 = error) and, on error, stores a malloc'ed NUL-terminated message in *err
 * which the caller releases with zkm_b200_free_string().
 *
 * Field elements are Goldilocks (p = 2^64 - 2^32 + 1) canonical u64, little endian — the memory image
 * of plonky2's `GoldilocksField(u64)` after `to_canonical_u64()`.  Extension elements are 2 u64.
 * Digests ("HashOut") are 4 u64.
 *
 * Threading: one context per process and device; calls are not re-entrant.
 */
#ifndef ZKM_B200_H
#define ZKM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One STARK table's trace: `PolynomialValues<F>` per column (reference prover.rs:133,
 * trace_poly_values: [Vec<PolynomialValues<F>>; NUM_TABLES]).  cols[c] points at 2^log_n u64. */
typedef struct {
    const uint64_t* const* cols;
    uint32_t ncols;
    uint32_t log_n;
} zkm_table_t;

/* StarkConfig / FriConfig (reference prover/src/config.rs:4-29). */
typedef struct {
    uint32_t rate_bits;        /* 2  */
    uint32_t cap_height;       /* 4  */
    uint32_t pow_bits;         /* 16 */
    uint32_t num_queries;      /* 37 */
    uint32_t num_challenges;   /* 2  */
    uint32_t arity_bits;       /* 4  (ConstantArityBits(4, 5)) */
    uint32_t final_poly_bits;  /* 5  */
} zkm_stark_config_t;

/* Opaque device-resident PolynomialBatch (coefficients + LDE + Merkle tree). */
typedef struct zkm_batch zkm_batch_t;

void zkm_b200_free_string(char* s);
void zkm_b200_free(void* p);

/* StarkConfig::standard_fast_config() (config.rs:17-29). */
void zkm_b200_standard_fast_config(zkm_stark_config_t* out);

/* Create the CUDA context, streams and twiddle tables on `device`.  Fails (no CPU fallback) when no
 * usable GPU is present. */
int zkm_b200_init(int device, char** err);
int zkm_b200_shutdown(char** err);
/* Number of CUDA kernels launched by this library since init (for bench.py's gpu_launches). */
uint64_t zkm_b200_launch_count(void);
/* Blocks until all device work queued by the library has finished. */
int zkm_b200_sync(char** err);

/* Worker contexts.  The library is not re-entrant on one context (reference callers are single-threaded at this level:
 * `&mut TimingTree`, prover.rs:130-140), but a host thread may bind a worker context -- its own streams, twiddle tables and
 * device-memory arena on the initialised device -- and every call it makes then runs there.  Two threads bound to two
 * workers can prove two segments at once: the latency-bound phases of one proof (small tables, transcript round trips)
 * overlap the throughput-bound kernels of the other.  Proofs do not depend on the context they were computed on. */
typedef struct zkm_worker zkm_worker_t;
int zkm_b200_worker_create(zkm_worker_t** out, char** err);
int zkm_b200_worker_bind(zkm_worker_t* w, char** err);      /* NULL: back to the process-wide context */
void zkm_b200_worker_destroy(zkm_worker_t* w);

/* In-segment sharding (BASELINE north_star: "partition independent trace-table NTTs and Merkle subtrees
 * across the GPUs ... NCCL only for the final cap/root gather").  world = 2, 4 or 8 processes, one per GPU, form a group: after
 * zkm_b200_shard_init every prove call on the process-wide context is COOPERATIVE -- all ranks of the group must make the same
 * call with the same inputs, and every rank returns the same proof, bit-identical to the single-GPU proof.  Each rank computes
 * the coset transforms, leaf hashes and Merkle subtrees of the LDE cosets it owns (4/world of the 4 cosets = 16/world whole cap
 * subtrees; with 8 ranks two ranks share a coset's transform and hash one half of its leaves each) and its share of the quotient; the exchanges are the cap entries of every commitment (ncclAllGather, 512 B per
 * tree), the two halves of the quotient values (ncclBroadcast) and the opened rows/paths of the FRI queries (ncclAllGather).
 * Rank 0 obtains the 128-byte NCCL unique id and the caller distributes it (bench.py / zkm_b200/multi.py: torch.distributed
 * broadcast).  NCCL is bound at run time (libnccl.so.2); single-GPU use does not need it.  world = 1 leaves sharding off. */
int zkm_b200_shard_unique_id(uint8_t out[128], char** err);
int zkm_b200_shard_init(int rank, int world, const uint8_t id[128], char** err);
int zkm_b200_shard_shutdown(char** err);

/* CUDA-event stopwatch on the library's stream (bench.py times the whole step with it). */
int zkm_b200_timer_start(char** err);
int zkm_b200_timer_stop(double* ms, char** err);
/* Device-side timing per kernel family (CUDA events on the launching stream) — the counterpart of the
 * reference's TimingTree scopes (prover.rs:86,144-167,191-215,...).  `bytes` is the algorithmic byte
 * count the family processed (DESIGN.md lists the per-unit figures).  families() returns a malloc'ed
 * '\n'-separated list, released with zkm_b200_free_string. */
void zkm_b200_profile_enable(int on);
int zkm_b200_profile_reset(char** err);
int zkm_b200_profile_get(const char* family, double* ms, uint64_t* launches, double* bytes, char** err);
/* Timings keyed by the reference's TimingTree scope strings (prover.rs:146 "compute all trace commitments", :152 "compute trace
 * commitment for {Table:?}", :204 "compute all proofs given commitments", :250-411 "prove {..} STARK", :513 "compute auxiliary
 * polynomials commitment", :545 "compute quotient polys", :578 "compute quotient commitment", :620 "compute openings proof";
 * the reference's :193 "compute CTL data" and :479 "compute lookup helper columns" are one scope per table here, and
 * "compute openings" brackets StarkOpeningSet::new).  When enabled, every prove call records CUDA events at the scope
 * boundaries; zkm_b200_last_timing returns the scopes of the calling thread's last proof as malloc'ed text, one line per scope
 * in opening order: "<depth>\t<milliseconds of device time>\t<scope>\n" (release with zkm_b200_free_string).  The Rust shim
 * logs them next to the TimingTree it was handed (a TimingTree cannot be given durations from outside). */
void zkm_b200_timing_enable(int on);
char* zkm_b200_last_timing(void);
/* Second per-family counter: bytes moved by the shared-memory passes for "ntt_pass" (16 B per element per pass), Poseidon
 * permutations for the hashing families ("leaf_hash", "merkle_levels", "leaf_hash_rows"), 0 elsewhere. */
int zkm_b200_profile_get_traffic(const char* family, double* aux, char** err);
char* zkm_b200_profile_families(void);

/* ---- the prover --------------------------------------------------------------------------------
 *
 * zkm_b200_prove_with_traces replaces reference prover/src/prover.rs:130-140
 *   prove_with_traces(all_stark, config, trace_poly_values: [Vec<PolynomialValues<F>>; NUM_TABLES],
 *                     public_values, timing) -> Result<AllProof<F, C, D>>
 * for F = GoldilocksField, C = PoseidonGoldilocksConfig, D = 2.  tables[] is indexed by the `Table`
 * enum (all_stark.rs:97-110); roots_before/roots_after/userdata are PublicValues (proof.rs:52-61).
 * Errors the reference raises as panics/ensure! come back as -1 with the same message text
 * ("FRI total reduction arity is too large.", "No CTL?", "Opening point is in the subgroup.",
 * "Non-binary filter?").
 *
 * Proof buffer layout (u64 words, little endian; extension element = 2 words, digest = 4 words,
 * "vec X" = length word then the items):
 *   magic "ZKMPROOF", version 1, num_tables,
 *   num_challenges, (beta, gamma) x num_challenges               -- AllProof.ctl_challenges
 *   roots_before[8], roots_after[8], vec userdata bytes           -- AllProof.public_values
 *   per table (StarkProofWithMetadata, proof.rs:178-201):
 *     init_challenger_state[12]
 *     vec digest trace_cap, vec digest auxiliary_polys_cap, vec digest quotient_polys_cap
 *     vec ext local_values, vec ext next_values, vec ext auxiliary_polys, vec ext auxiliary_polys_next,
 *     vec F ctl_zs_first, vec ext quotient_polys                  -- StarkOpeningSet (proof.rs:283-296)
 *     FriProof: vec (vec digest) commit_phase_merkle_caps,
 *               vec query_round { vec oracle { vec F leaf row, vec digest siblings },
 *                                 vec step   { vec ext evals,  vec digest siblings } },
 *               vec ext final_poly, pow_witness
 * The buffer is malloc'ed by the library and released with zkm_b200_free. */
int zkm_b200_prove_with_traces(const zkm_table_t tables[12], const uint32_t roots_before[8], const uint32_t roots_after[8],
                               const uint8_t* userdata, uint32_t userdata_len, const zkm_stark_config_t* cfg,
                               uint64_t** proof_out, size_t* proof_words, char** err);
/* A table as its generator leaves it: 2^log_n rows of ncols words, row-major -- the `Vec<[F; COLUMNS]>` that
 * trace_rows_to_poly_values (reference util.rs:37-47) transposes on the CPU in Traces::into_tables
 * (witness/traces.rs:274-305).  Passing the rows moves that transposition onto the device. */
typedef struct {
    const uint64_t* rows;
    uint32_t ncols;
    uint32_t log_n;
} zkm_table_rows_t;
/* prove_with_traces where table t is taken from row_tables[t] when row_tables[t].rows != NULL and from tables[t]
 * otherwise.  Ten tables come out of their generators as finished rows.  The Arithmetic table (index 0) passed as rows is
 * taken as ArithmeticStark::generate_trace has it BEFORE generate_range_checks (arithmetic_stark.rs:155-192: operation rows
 * + zero padding to >= 2^16 rows): RANGE_COUNTER and RC_FREQUENCIES are then generated on the device (:127-153), and a
 * shared cell >= 2^16 is reported as the reference's assertion text.  Memory is sorted column-wise upstream: pass columns. */
int zkm_b200_prove_with_trace_rows(const zkm_table_t tables[12], const zkm_table_rows_t row_tables[12],
                                   const uint32_t roots_before[8], const uint32_t roots_after[8], const uint8_t* userdata,
                                   uint32_t userdata_len, const zkm_stark_config_t* cfg, uint64_t** proof_out,
                                   size_t* proof_words, char** err);
/* The Memory table from the log of memory operations (reference memory/memory_stark.rs:133-244
 * MemoryStark::generate_trace: stable sort by (context, segment, virt, timestamp), fill_gaps, padding, first-change flags,
 * range check, counter, frequencies; writes to register 0 recorded as 0).  ops = n_ops x 7 words in push order:
 * context, segment, virt, timestamp, is_read, value, filter.  Returns 13 columns (column-major, 2^*log_n_out rows each) in a
 * malloc'ed buffer released with zkm_b200_free -- the Memory entry of tables[] for the calls above. */
int zkm_b200_memory_trace(const uint64_t* ops, size_t n_ops, uint64_t** cols_out, uint32_t* log_n_out, char** err);
/* prove_with_trace_rows (row_tables may be NULL: all tables column-major) where the Memory table (index 11) is generated on
 * the device from the operation log and never crosses PCIe as a table; tables[11] / row_tables[11] are ignored. */
int zkm_b200_prove_with_memory_ops(const zkm_table_t tables[12], const zkm_table_rows_t* row_tables, const uint64_t* memory_ops,
                                   size_t n_memory_ops, const uint32_t roots_before[8], const uint32_t roots_after[8],
                                   const uint8_t* userdata, uint32_t userdata_len, const zkm_stark_config_t* cfg,
                                   uint64_t** proof_out, size_t* proof_words, char** err);
/* Tables generated on the device from their operation logs (the per-table generators that
 * Traces::into_tables runs on the CPU, witness/traces.rs:271-301).  Log formats, one entry per operation, u64 words:
 *   Arithmetic (table 0) 3 words: operator (the IS_* column index 0..25 = BinaryOperator::row_filter, arithmetic/mod.rs:135-165),
 *                                 input0, input1 (u32) as Operation::binary receives them       -> ArithmeticStark::generate_trace
 *                                 (arithmetic_stark.rs:155-192: one or two rows per operation, range-check columns included)
 *   Memory   (table 11)  7 words: context, segment, virt, timestamp, is_read, value, filter      -> MemoryStark::generate_trace
 *   Logic    (table 10)  3 words: operator (0 AND, 1 OR, 2 XOR, 3 NOR), input0, input1 (u32)     -> LogicStark::generate_trace (logic.rs:108-183)
 *   Poseidon (table 2)  13 words: the 12 input elements (canonical), timestamp                   -> PoseidonStark::generate_trace
 *                                                                                                   (poseidon_stark.rs:51-145)
 *   Keccak   (table 4)  26 words: the 25 input lanes (input[y * 5 + x] = lane (x, y)), timestamp -> KeccakStark::generate_trace
 *                                                                                                   (keccak_stark.rs:62-237, 24 rows per permutation)
 *   ShaExtend (table 6)  5 words: w[i-15], w[i-2], w[i-16], w[i-7] (u32), timestamp                -> ShaExtendStark::generate_trace
 *                                                                                                   (sha_extend_stark.rs:122-237, one row per entry)
 *   ShaExtendSponge (7) 13 words: round i (0..47), the same four words, input_virt[4], output_virt, context, segment, timestamp
 *                                                                                                -> sha_extend_sponge_stark.rs:128-227
 *   ShaCompress (8)     15 words: a..h (u32), w_i, k_i, round (0..64), w_i_virt, segment, context, timestamp -- one entry per
 *                                 ROW, the reference's ([u8; 41], MemoryAddress, usize)           -> sha_compress_stark.rs:234-391
 *   ShaCompressSponge (9) 86 words: hx[8], w[64] (u32), hx_virt[8], w_start virt / segment / context, context, segment,
 *                                 timestamp                                                       -> sha_compress_sponge_stark.rs:120-237
 *   PoseidonSponge (3), KeccakSponge (5): a VARIABLE-width log.  Word 0 = the number of words of the whole log; then per operation
 *                                 context, segment, timestamp, len (bytes), n_addr, virt[n_addr], the input bytes packed little endian
 *                                 8 per word ({Keccak,Poseidon}SpongeOp {base_address, timestamp, input}: context / segment of
 *                                 base_address[0], virt[k] = base_address[k].virt); len / rate + 1 rows per operation
 *                                 (keccak_sponge_stark.rs:222-447 rate 136, poseidon_sponge_stark.rs:187-365 rate 32); n_ops = operations
 * The Cpu table (1) has no log: its rows ARE the interpreter's output (pass them row-major, zkm_b200_prove_with_trace_rows).
 * Heights follow the reference: next power of two of max(number of rows, min_rows), min_rows = max(2^cap_height, 64).
 * zkm_b200_table_from_ops returns the finished columns (column-major, malloc'ed; zkm_b200_free); zkm_b200_prove_with_ops is
 * prove_with_trace_rows where every table t with op_logs[t].ops != NULL is generated on the device instead of being read from
 * tables[t] / row_tables[t] -- only the log crosses PCIe (Logic: 24 B instead of 552 B per row, Arithmetic: 24 B instead of 432 B). */
typedef struct {
    const uint64_t* ops;
    size_t n_ops;
} zkm_op_log_t;
int zkm_b200_table_from_ops(uint32_t table, const uint64_t* ops, size_t n_ops, uint32_t min_rows, uint64_t** cols_out,
                            uint32_t* log_n_out, char** err);
int zkm_b200_prove_with_ops(const zkm_table_t tables[12], const zkm_table_rows_t* row_tables, const zkm_op_log_t op_logs[12],
                            const uint32_t roots_before[8], const uint32_t roots_after[8], const uint8_t* userdata,
                            uint32_t userdata_len, const zkm_stark_config_t* cfg, uint64_t** proof_out, size_t* proof_words,
                            char** err);
/* Same prover over another System of tables (zkm_b200/csrc/tables/systems.h: 0 = AllStark, 1 = Logic,
 * 2 = Poseidon+Logic+Memory, 3 = Poseidon, 4 = Memory, 5 = Arithmetic, 6 = Keccak+KeccakSponge+Logic+Memory,
 * 7 = Poseidon+PoseidonSponge+Memory, 8 = ShaExtend+ShaExtendSponge+Logic+Memory, 9 = ShaCompress+ShaCompressSponge+
 * Logic+Memory, 10 = Cpu+Arithmetic+Logic+Memory): slices of AllStark with its real cross-table lookups, whose valid
 * traces the tests generate from restated witness generators, for prove -> verify parity tests. */
int zkm_b200_prove_system(int system_id, const zkm_table_t* tables, uint32_t num_tables, const uint32_t* roots_before,
                          const uint32_t* roots_after, const uint8_t* userdata, uint32_t userdata_len,
                          const zkm_stark_config_t* cfg, uint64_t** proof_out, size_t* proof_words, char** err);
/* As above with the trace columns already resident in device memory: d_tables[t] is a device pointer to
 * shapes[t].ncols * 2^log_n u64, column-major (shapes[t].cols is ignored).  bench.py's `value` metric. */
int zkm_b200_prove_system_device(int system_id, const zkm_table_t* shapes, const uint64_t* const* d_tables, uint32_t num_tables,
                                 const uint32_t* roots_before, const uint32_t* roots_after, const uint8_t* userdata,
                                 uint32_t userdata_len, const zkm_stark_config_t* cfg, uint64_t** proof_out, size_t* proof_words,
                                 char** err);

/* ---- proof wire format -------------------------------------------------------------------------
 *
 * zkm_b200_proof_table_json writes what `serde_json::to_string(&all_proof.stark_proofs[table].proof)` gives for the reference's
 * `StarkProof<F, C, D>` (proof.rs:177-189, #[derive(Serialize)]; the text whose length prove_segments logs as "proof size",
 * prover/examples/utils/src/utils.rs:156-161): compact JSON, fields in declaration order -- trace_cap, auxiliary_polys_cap,
 * quotient_polys_cap, openings {local_values, next_values, auxiliary_polys, auxiliary_polys_next, ctl_zs_first,
 * quotient_polys}, opening_proof {commit_phase_merkle_caps, query_round_proofs [{initial_trees_proof {evals_proofs}, steps
 * [{evals, merkle_proof {siblings}}]}], final_poly {coeffs}, pow_witness}; field elements are decimal u64, extension elements
 * [a, b], digests {"elements":[..4..]}.  zkm_b200_public_values_json does the same for `PublicValues` (proof.rs:52-66; the
 * file recursion/src/lib.rs:142-146 writes).  Host-only: no device, no zkm_b200_init.  Strings are malloc'ed (zkm_b200_free_string). */
int zkm_b200_proof_table_json(const uint64_t* proof, size_t proof_words, uint32_t table, char** json_out, size_t* json_len, char** err);
int zkm_b200_public_values_json(const uint64_t* proof, size_t proof_words, char** json_out, size_t* json_len, char** err);

/* The segment file: `serde_json::to_vec(&segment)` for the emulator's `Segment` (emulator/src/state.rs:33-48; written by
 * split_segment :1498-1503, read back by prove_segments).  mem_image is given as the pages the segment read (Memory::rtrace,
 * memory.rs:524-538 get_input_image): n_pages pages of 4096 bytes with strictly ascending page indices (address >> 12); the
 * text lists every word, key = address (decimal string, BTreeMap order), value = u32::from_le_bytes of the page bytes.
 * Host-only.  The string is malloc'ed (zkm_b200_free_string). */
typedef struct {
    const uint32_t* page_indices; const uint8_t* pages; size_t n_pages;
    uint32_t pc, segment_id;
    uint8_t pre_image_id[32], pre_hash_root[32], image_id[32], page_hash_root[32];
    uint32_t end_pc;
    uint64_t step;
    const uint8_t* const* input_stream; const size_t* input_stream_lens; size_t n_input_streams;
    uint64_t input_stream_ptr;
    const uint8_t* public_values_stream; size_t public_values_stream_len;
    uint64_t public_values_stream_ptr;
} zkm_segment_t;
int zkm_b200_segment_json(const zkm_segment_t* seg, char** json_out, size_t* json_len, char** err);

/* ---- emulator segment splitter: page hashes and image id ---------------------------------------
 *
 * At every segment boundary the reference's InstrumentedState::split_segment (emulator/src/state.rs:1460-1530) calls
 * Memory::update_page_hash (emulator/src/memory.rs:415-436) -- hash_page (:81-89, 129 Poseidon permutations) of every 4 KiB
 * page the segment wrote, the digests stored into the L1 hash pages, those re-hashed into L2 and the root page -- and
 * Memory::compute_image_id (:438-471).  The pages are independent: these entry points do the hashing on the device.
 *
 * zkm_b200_hash_pages: hash_page of n_pages contiguous pages -> n_pages x 32 bytes (the four digest words, little endian).
 * A zkm_pagetree_t holds the hash pages at and above MAX_MEMORY = 0x80000000 (absent pages have the reference's
 * CONST_HASH_PAGES content, :91-125).  zkm_b200_pagetree_split(tree, indices, pages, n, registers, pc) is update_page_hash
 * over the dirty main-memory pages `pages[k]` (page index indices[k] = address >> 12, < 0x80000; the reference's wtrace[0])
 * followed by compute_image_id(pc, registers = State::get_registers_bytes(), 39 x 4 bytes): image_id_out and
 * page_hash_root_out receive 32 bytes each.  "compute image ID fail" is returned when no page was ever hashed (the
 * reference's panic).  zkm_b200_pagetree_page reads one hash page back (they are part of the segment's memory image). */
typedef struct zkm_pagetree zkm_pagetree_t;
int zkm_b200_hash_pages(const uint8_t* pages, size_t n_pages, uint8_t* digests_out, char** err);
int zkm_b200_pagetree_create(zkm_pagetree_t** out, char** err);
void zkm_b200_pagetree_destroy(zkm_pagetree_t* t);
int zkm_b200_pagetree_split(zkm_pagetree_t* t, const uint32_t* page_indices, const uint8_t* pages, size_t n_pages,
                            const uint8_t* registers, uint32_t pc, uint8_t* image_id_out, uint8_t* page_hash_root_out, char** err);
int zkm_b200_pagetree_page(const zkm_pagetree_t* t, uint32_t page_index, uint8_t* out, int* present, char** err);
/* Seeds one hash page (index 0x80000 ..= 0x81020) from the caller's memory: an emulator state resumed from a segment file
 * (State::load_seg, emulator/src/state.rs:141-190, as split_seg_into_segs does, utils.rs:62-109) carries its hash pages in the
 * memory image.  Host-only: needs no device. */
int zkm_b200_pagetree_set_page(zkm_pagetree_t* t, uint32_t page_index, const uint8_t* data, char** err);

/* The whole of InstrumentedState::split_segment (emulator/src/state.rs:1477-1530) except the step loop that decides WHEN to split:
 * a zkm_splitter_t owns the hash pages and the `pre_*` bookkeeping of InstrumentedState (:556-596; all zero at creation).
 * zkm_b200_splitter_split(s, state, proof, ..):
 *   1. update_page_hash over state->dirty_pages (Memory::wtrace[0]) and compute_image_id(state->pc, state->registers) on the device;
 *   2. if `proof`: the Segment { mem_image = state->read_pages (Memory::rtrace, i.e. get_input_image()), segment_id = pre_segment_id,
 *      pc = pre_pc, pre_hash_root, pre_image_id, image_id, end_pc = state->pc, step = state->step, page_hash_root, input_stream =
 *      pre_input, input_stream_ptr = pre_input_ptr, public_values_stream = pre_public_values, .. } as its serde_json text in
 *      *segment_json_out (malloc'ed; zkm_b200_free_string) -- what the reference writes to "{output}/{segment_id}" -- and
 *      pre_segment_id += 1;
 *   3. pre_input / pre_public_values / pre_pc / pre_image_id / pre_hash_root take the current values.
 * split_prog_into_segs (emulator/src/utils.rs:23-57) calls it once with proof = 0 before the first step and with proof = 1 at
 * every boundary and at exit.  Page indices must be strictly ascending (BTreeMap order). */
typedef struct zkm_splitter zkm_splitter_t;
typedef struct {
    const uint32_t* dirty_page_indices; const uint8_t* dirty_pages; size_t n_dirty_pages;      /* wtrace[0] */
    const uint32_t* read_page_indices; const uint8_t* read_pages; size_t n_read_pages;         /* rtrace */
    const uint8_t* registers;                                                                   /* get_registers_bytes(): 156 B */
    uint32_t pc;
    uint64_t step;
    const uint8_t* const* input_stream; const size_t* input_stream_lens; size_t n_input_streams;
    uint64_t input_stream_ptr;
    const uint8_t* public_values_stream; size_t public_values_stream_len;
    uint64_t public_values_stream_ptr;
} zkm_split_state_t;
int zkm_b200_splitter_create(zkm_splitter_t** out, char** err);
void zkm_b200_splitter_destroy(zkm_splitter_t* s);
zkm_pagetree_t* zkm_b200_splitter_pagetree(zkm_splitter_t* s);
uint32_t zkm_b200_splitter_segment_count(const zkm_splitter_t* s);
int zkm_b200_splitter_split(zkm_splitter_t* s, const zkm_split_state_t* state, int proof, char** segment_json_out, size_t* segment_json_len,
                            uint8_t* image_id_out, uint8_t* page_hash_root_out, char** err);

/* ---- column-layout handshake ---------------------------------------------------------------------------------------
 *
 * The constraint kernels address trace columns by index.  Those indices are the memory layout of the reference's column
 * structs: `CpuColumnsView` is #[repr(C)] (cpu/columns/mod.rs:68-118), but its 102-wide `general` union overlays nine
 * views that are NOT #[repr(C)] (cpu/columns/general.rs:8-18,143-201), so their field order is formally up to rustc.  The Rust
 * shim therefore reports where its compiler actually put every field this library reads -- (key, value) pairs taken from
 * `COL_MAP` (mod.rs:184-189) and from `NUM_*_COLUMNS` -- and zkm_b200_layout_check compares them with the constants the
 * kernels were compiled with (the headers under zkm_b200/csrc/tables).  A mismatch is an error naming the key, before any proof is made.
 * Keys: ZKM_LK_NUM_COLUMNS + t = number of trace columns of table t (Table enum order, all_stark.rs:97-110); the CPU keys
 * below are absolute column indices, except the *_REL keys, which are relative to ZKM_LK_CPU_GENERAL resp. to the start of
 * one memory channel. */
typedef enum {
    ZKM_LK_NUM_COLUMNS = 0,                 /* + table index 0..11 */
    ZKM_LK_CPU_IS_BOOTSTRAP_KERNEL = 100, ZKM_LK_CPU_IS_EXIT_KERNEL, ZKM_LK_CPU_CONTEXT, ZKM_LK_CPU_CODE_CONTEXT,
    ZKM_LK_CPU_PROGRAM_COUNTER, ZKM_LK_CPU_NEXT_PROGRAM_COUNTER, ZKM_LK_CPU_IS_KERNEL_MODE,
    ZKM_LK_CPU_OP_BINARY_OP /* first op flag */, ZKM_LK_CPU_OP_SYSCALL /* last op flag */,
    ZKM_LK_CPU_BRANCH_SHOULD_JUMP, ZKM_LK_CPU_BRANCH_IS_NE,
    ZKM_LK_CPU_OPCODE_BITS, ZKM_LK_CPU_RS_BITS, ZKM_LK_CPU_RT_BITS, ZKM_LK_CPU_RD_BITS, ZKM_LK_CPU_SHAMT_BITS, ZKM_LK_CPU_FUNC_BITS,
    ZKM_LK_CPU_IS_POSEIDON_SPONGE, ZKM_LK_CPU_IS_KECCAK_SPONGE, ZKM_LK_CPU_IS_SHA_EXTEND_SPONGE, ZKM_LK_CPU_IS_SHA_COMPRESS_SPONGE,
    ZKM_LK_CPU_GENERAL, ZKM_LK_CPU_MEMIO_IS_LH, ZKM_LK_CPU_MEMIO_AUX_FILTER, ZKM_LK_CPU_CLOCK, ZKM_LK_CPU_MEM_CHANNELS,
    ZKM_LK_CPU_MEM_CHANNEL_STRIDE,          /* columns per MemoryChannelView */
    ZKM_LK_CPU_CH_USED_REL, ZKM_LK_CPU_CH_IS_READ_REL, ZKM_LK_CPU_CH_ADDR_CONTEXT_REL, ZKM_LK_CPU_CH_ADDR_SEGMENT_REL,
    ZKM_LK_CPU_CH_ADDR_VIRTUAL_REL, ZKM_LK_CPU_CH_VALUE_REL,
    ZKM_LK_CPU_G_SYSCALL_COND_REL = 200, ZKM_LK_CPU_G_SYSCALL_SYSNUM_REL, ZKM_LK_CPU_G_SYSCALL_A0_REL, ZKM_LK_CPU_G_SYSCALL_A1_REL,
    ZKM_LK_CPU_G_MISC_RS_BITS_REL, ZKM_LK_CPU_G_MISC_IS_MSB_REL, ZKM_LK_CPU_G_MISC_IS_LSB_REL, ZKM_LK_CPU_G_MISC_AUXM_REL,
    ZKM_LK_CPU_G_MISC_AUXL_REL, ZKM_LK_CPU_G_MISC_AUXS_REL, ZKM_LK_CPU_G_MISC_RD_INDEX_REL, ZKM_LK_CPU_G_MISC_RD_INDEX_EQ_0_REL,
    ZKM_LK_CPU_G_MISC_RD_INDEX_EQ_29_REL,
    ZKM_LK_CPU_G_IO_RS_LE_REL, ZKM_LK_CPU_G_IO_RT_LE_REL, ZKM_LK_CPU_G_IO_MEM_LE_REL, ZKM_LK_CPU_G_IO_AUX_RS0_MUL_RS1_REL,
    ZKM_LK_CPU_G_LOGIC_DIFF_PINV_REL, ZKM_LK_CPU_G_HASH_VALUE_REL, ZKM_LK_CPU_G_KHASH_VALUE_REL, ZKM_LK_CPU_G_SHASH_VALUE_REL,
    ZKM_LK_CPU_G_ELEMENT_VALUE_REL
} zkm_layout_key_t;
/* pairs = n_pairs x (key, value).  Needs no device and no zkm_b200_init.  Returns -1 with a message naming the first
 * mismatching key (or an unknown key). */
int zkm_b200_layout_check(const uint32_t* pairs, size_t n_pairs, char** err);
/* The library's own view: writes up to max_pairs (key, value) pairs for every key above, returns their number in *n_pairs. */
int zkm_b200_layout_describe(uint32_t* pairs, size_t max_pairs, size_t* n_pairs, char** err);

/* ---- staged API (stage-by-stage parity against the oracle) ------------------------------- */

/* PolynomialBatch::from_values(values, rate_bits, blinding=false, cap_height) — prover.rs:154-163,
 * 514-521.  Host columns in; cap_out receives (1<<cap_height)*4 u64. */
int zkm_b200_commit_values(const zkm_table_t* table, uint32_t rate_bits, uint32_t cap_height,
                           zkm_batch_t** out, uint64_t* cap_out, char** err);
/* PolynomialBatch::from_coeffs — prover.rs:576-587. */
int zkm_b200_commit_coeffs(const zkm_table_t* table, uint32_t rate_bits, uint32_t cap_height,
                           zkm_batch_t** out, uint64_t* cap_out, char** err);
/* Same as commit_values but the columns are already in device memory: `d_values` is a device pointer
 * to ncols*2^log_n u64, column-major.  Used by bench.py for the HBM-resident `value` metric. */
int zkm_b200_commit_values_device(const uint64_t* d_values, uint32_t ncols, uint32_t log_n,
                                  uint32_t rate_bits, uint32_t cap_height, zkm_batch_t** out,
                                  uint64_t* cap_out, char** err);
/* Fills a device buffer (ncols x 2^log_n u64, column-major) with the synthetic SplitMix64 columns of
 * BASELINE.md §3 (stream seed | column index, value mod p).  Benchmark input generator. */
int zkm_b200_synth_columns_device(uint64_t* d_out, uint32_t ncols, uint32_t log_n, uint64_t seed, char** err);
/* Synthetic trace of table `table` of System `system_id`: uniform cells, with the columns
 * read by CTL filters drawn so that every filter is 0/1.  _device writes ncols*2^log_n u64 (column-major)
 * to a device pointer, the other variant to a host buffer.  Benchmark / parity input generator. */
int zkm_b200_synth_trace_device(int system_id, uint32_t table, uint32_t log_n, uint64_t seed, uint64_t* d_out, char** err);
int zkm_b200_synth_trace(int system_id, uint32_t table, uint32_t log_n, uint64_t seed, uint64_t* host_out, char** err);
/* Number of tables of a System and the column count of each (AllStark: 12 tables). */
int zkm_b200_system_shape(int system_id, uint32_t* num_tables, uint32_t* ncols_out, uint32_t max_tables, char** err);
void zkm_b200_batch_free(zkm_batch_t* b);
/* Coefficients of polynomial `col` (2^log_n u64). */
int zkm_b200_batch_get_coeffs(const zkm_batch_t* b, uint32_t col, uint64_t* out, char** err);
/* LDE values of polynomial `col` on 7*H_{4n} in natural order (2^(log_n+rate_bits) u64). */
int zkm_b200_batch_get_lde(const zkm_batch_t* b, uint32_t col, uint64_t* out, char** err);
/* MerkleTree::get(leaf) + MerkleTree::prove(leaf): leaf_out gets ncols u64, siblings_out gets
 * (log_n + rate_bits - cap_height)*4 u64, leaf level first. */
int zkm_b200_batch_open(const zkm_batch_t* b, uint32_t leaf_index, uint64_t* leaf_out,
                        uint64_t* siblings_out, char** err);

/* The finer-grained seam of SURVEY section 8(b) for ONE table of a System, under caller-given challenges (no transcript), for
 * stage-level parity against the oracle:
 *   aux_out       num_aux x 2^log_n words, column-major: lookup_helper_columns (lookup.rs:46-124) then the CTL helper and Z
 *                 columns of cross_table_lookup_data (cross_table_lookup.rs:634-703,801-872), in the order prove_single_table
 *                 commits them (prover.rs:469-522)
 *   quotient_out  num_challenges x 2^(log_n + 1) coefficients: compute_quotient_polys + coset_ifft (prover.rs:645-789); chunk k
 *                 of polynomial a (2^log_n coefficients) is quotient polynomial 2a + k (:560-587)
 *   openings_out  StarkOpeningSet::new (proof.rs:299-334): local_values[C], next_values[C], auxiliary_polys[num_aux],
 *                 auxiliary_polys_next[num_aux] (2 words each), ctl_zs_first (1 word each), quotient_polys[2 num_challenges]
 * ctl_challenges = num_challenges x (beta, gamma), alphas = num_challenges words, zeta = 2 words.  Buffers are malloc'ed
 * (zkm_b200_free). */
int zkm_b200_stage_table(int system_id, uint32_t table_index, const zkm_table_t* table, const zkm_stark_config_t* cfg,
                         const uint64_t* ctl_challenges, const uint64_t* alphas, const uint64_t* zeta, uint64_t** aux_out,
                         uint32_t* num_aux_out, uint64_t** quotient_out, uint64_t** openings_out, size_t* openings_words, char** err);

/* Raw transforms on host buffers (column-major, ncols x 2^log_n), for NTT parity tests.
 * kind: 0 = fft, 1 = ifft, 2 = coset_ifft(7). */
int zkm_b200_ntt(uint64_t* data, uint32_t ncols, uint32_t log_n, int kind, char** err);
/* Poseidon permutations on `count` independent 12-word states (host buffer), run on the GPU. */
int zkm_b200_poseidon_permute(uint64_t* states, size_t count, char** err);
/* The permutation of the prover's host-side Fiat-Shamir transcript (plonky2 iop/challenger.rs duplexing as reached from
 * reference prover/src/prover.rs:182-190,466,524-527,588-591,610), on `count` independent 12-word states in place.
 * Runs on the CPU, needs no device and no zkm_b200_init: exported so the transcript's hash can be checked on its own. */
int zkm_b200_transcript_permute(uint64_t* states, size_t count, char** err);

#ifdef __cplusplus
}
#endif
#endif /* ZKM_B200_H */
