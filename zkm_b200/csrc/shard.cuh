// In-segment sharding: one STARK proof computed by G = 2, 4 or 8 GPUs (one process per GPU), SURVEY section 8(e).
//
// The 4n-point LDE domain 7*H_{4n} is the disjoint union of the four cosets 7 w_{4n}^j H_n, and this library already stores
// every LDE coset-major (ntt.cuh).  Coset j is exactly the quarter bitrev2(j) of the bit-reversed leaf order, i.e. four whole
// subtrees of the 16-entry Merkle cap, so a rank that owns cosets [4r/G, 4(r+1)/G) computes those coset transforms, hashes
// those leaves and builds those subtrees without ever needing another rank's data.  The exchanges are:
//   * the cap entries of every sharded commitment (512 B per tree)              -- ncclAllGather
//   * the two halves of the quotient values (cosets 0 and 2 of the quotient domain; n x 2 words each) -- ncclBroadcast
//   * the opened leaf rows + authentication paths of the 37 FRI queries         -- ncclAllGather, owner's copy selected
// Everything else (auxiliary columns, openings, the FRI commit phase, the transcript) is computed identically on every
// rank from identical inputs, so all ranks observe the same caps, draw the same challenges and assemble the same proof:
// no challenge broadcast is needed and the proofs of all ranks are bit-identical to the single-GPU proof.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the library torch.distributed already loaded in the calling process):
// the single-GPU build has no link-time dependency on it.
#pragma once
#include "dev.cuh"

namespace zkm {

struct Shard {
    int rank = 0, world = 1;
    bool active() const { return world > 1; }
    // cosets (of the 4 = 2^rate_bits) this rank computes.  Up to 4 ranks: 4/world whole cosets each.  8 ranks: two ranks share
    // a coset -- both run its transform (the NTT is 20 % of a commitment, the hashing 80 %), each hashes one half of its
    // leaves (the rows i = part mod 2 of the coset, i.e. a contiguous eighth of the bit-reversed leaf order = 2 cap subtrees).
    int parts() const { return world > 4 ? world / 4 : 1; }           // ranks per coset
    int log_parts() const { return world > 4 ? 1 : 0; }
    int part() const { return rank % parts(); }
    int coset_begin() const { return rank * 4 / world; }
    int coset_count() const { return world >= 4 ? 1 : 4 / world; }
    bool owns_coset(int j) const { return j >= coset_begin() && j < coset_begin() + coset_count(); }
    // leaf segments: the leaves split into 4 * parts() contiguous segments; coset j, part p is segment bitrev2(j) * parts() + p
    int log_segs() const { return 2 + log_parts(); }
    int num_owned_segs() const { return coset_count(); }
    static int rank_of(int coset, int part, int world) { return world > 4 ? coset * (world / 4) + part : coset * world / 4; }
    static int coset_owner(int j, int world) { return rank_of(j, 0, world); }
};
inline int bitrev2(int j) { return ((j & 1) << 1) | (j >> 1); }
inline int shard_seg_of(int coset, int part, int world) { return world > 4 ? bitrev2(coset) * (world / 4) + part : bitrev2(coset); }

const Shard& shard();
// 128-byte NCCL unique id (rank 0 creates it, the caller distributes it to the other ranks out of band).
void shard_unique_id(unsigned char out[128]);
void shard_init(int rank, int world, const unsigned char id[128]);
void shard_shutdown();

// Collectives on the library stream (element type u64).
void shard_all_gather(const u64* d_send, u64* d_recv, size_t count_per_rank, cudaStream_t s);
void shard_broadcast(u64* d_buf, size_t count, int root, cudaStream_t s);

}  // namespace zkm
