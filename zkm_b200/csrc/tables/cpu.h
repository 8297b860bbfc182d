// CPU table (259 columns).  Column map: reference prover/src/cpu/columns/mod.rs:68-118 (CpuColumnsView,
// #[repr(C)]), columns/ops.rs:9-46 (33 op flags), columns/general.rs:8-18,143-201 (102-wide union of
// views; every view starts at offset 0 of the union).  Constraints in the reference's emission order
// (cpu_stark.rs:260-285): bootstrap_kernel.rs:308-351 -> decode.rs:66-95 -> jumps.rs:17-240 (jump/jumpi),
// :243-672 (branch) -> membus.rs:34-46 -> memio.rs:175-735 (load), :738-1214 (store) -> shift.rs:11-76 ->
// count.rs:10-70 -> syscall.rs:12-222 -> bits.rs:9-62 -> misc.rs:10-819 (rdhwr, condmov, teq, extract,
// ror, insert, maddu).  exit_kernel and contextops are commented out in the reference (:274,284).
// CTL column selectors: cpu_stark.rs:25-244.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace cpu {

constexpr int NUM_GP_CHANNELS = 9, NUM_CHANNELS = 10;      // membus.rs:10-20
constexpr int SEG_CODE = 0, SEG_SHIFT_TABLE = 3, SEG_REGISTER_FILE = 4;      // memory/segments.rs:4-16

enum {
    IS_BOOTSTRAP_KERNEL = 0, IS_EXIT_KERNEL, CONTEXT, CODE_CONTEXT, PROGRAM_COUNTER, NEXT_PROGRAM_COUNTER, IS_KERNEL_MODE,
    // OpsColumnsView
    OP_START,
    OP_BINARY_OP = OP_START, OP_BINARY_IMM_OP, OP_EQ_ISZERO, OP_LOGIC_OP, OP_LOGIC_IMM_OP, OP_MOVZ_OP, OP_MOVN_OP, OP_CLZ_OP, OP_CLO_OP,
    OP_SHIFT, OP_SHIFT_IMM, OP_KECCAK_GENERAL, OP_JUMPS, OP_JUMPI, OP_JUMPDIRECT, OP_BRANCH, OP_PC, OP_GET_CONTEXT, OP_SET_CONTEXT,
    OP_EXIT_KERNEL, OP_M_OP_LOAD, OP_M_OP_STORE, OP_NOP, OP_EXT, OP_INS, OP_MADDU, OP_RDHWR, OP_SIGNEXT8, OP_SIGNEXT16, OP_SWAPHALF,
    OP_TEQ, OP_ROR, OP_SYSCALL, OP_END,
    // CpuBranchView
    BR_SHOULD_JUMP = OP_END, BR_GT, BR_LT, BR_EQ, BR_IS_GT, BR_IS_LT, BR_IS_EQ, BR_IS_GE, BR_IS_LE, BR_IS_NE,
    OPCODE_BITS, RS_BITS = OPCODE_BITS + 6, RT_BITS = RS_BITS + 5, RD_BITS = RT_BITS + 5, SHAMT_BITS = RD_BITS + 5,
    FUNC_BITS = SHAMT_BITS + 5,
    IS_POSEIDON_SPONGE = FUNC_BITS + 6, IS_KECCAK_SPONGE, IS_SHA_EXTEND_SPONGE, IS_SHA_COMPRESS_SPONGE,
    GENERAL,
    MEMIO = GENERAL + 102,
    MEMIO_IS_LH = MEMIO, MEMIO_IS_LWL, MEMIO_IS_LW, MEMIO_IS_LBU, MEMIO_IS_LHU, MEMIO_IS_LWR, MEMIO_IS_SB, MEMIO_IS_SH, MEMIO_IS_SWL,
    MEMIO_IS_SW, MEMIO_IS_SWR, MEMIO_IS_LL, MEMIO_IS_SC, MEMIO_IS_SDC1, MEMIO_IS_LB, MEMIO_AUX_FILTER,
    CLOCK,
    MEM_CHANNELS,
    NUM_COLUMNS = MEM_CHANNELS + 6 * NUM_GP_CHANNELS
};
static_assert(OP_END - OP_START == 33 && GENERAL == 86 && MEMIO == 188 && CLOCK == 204 && NUM_COLUMNS == 259, "cpu column layout");
// MemoryChannelView
constexpr int CH_USED = 0, CH_IS_READ = 1, CH_ADDR_CONTEXT = 2, CH_ADDR_SEGMENT = 3, CH_ADDR_VIRTUAL = 4, CH_VALUE = 5;
ZKM_HD constexpr int ch(int channel, int field) { return MEM_CHANNELS + 6 * channel + field; }
// general-column views (offsets inside the union)
constexpr int G_SYSCALL_COND = GENERAL, G_SYSCALL_SYSNUM = GENERAL + 12, G_SYSCALL_A0 = GENERAL + 24, G_SYSCALL_A1 = GENERAL + 27;
constexpr int G_MISC_RS_BITS = GENERAL, G_MISC_IS_MSB = GENERAL + 32, G_MISC_IS_LSB = GENERAL + 64, G_MISC_AUXM = GENERAL + 96,
              G_MISC_AUXL = GENERAL + 97, G_MISC_AUXS = GENERAL + 98, G_MISC_RD_INDEX = GENERAL + 99, G_MISC_RD_INDEX_EQ_0 = GENERAL + 100,
              G_MISC_RD_INDEX_EQ_29 = GENERAL + 101;
constexpr int G_LOGIC_DIFF_PINV = GENERAL;
constexpr int G_IO_RS_LE = GENERAL, G_IO_RT_LE = GENERAL + 32, G_IO_MEM_LE = GENERAL + 64, G_IO_AUX_RS0_MUL_RS1 = GENERAL + 96;
constexpr int G_HASH_VALUE = GENERAL, G_KHASH_VALUE = GENERAL, G_SHASH_VALUE = GENERAL, G_ELEMENT_VALUE = GENERAL;

constexpr uint64_t GOLDILOCKS_INVERSE_2EXP32 = 18446744065119617026ull;      // jumps.rs:15
constexpr uint64_t MIPSEBADF = 0x9;                                           // witness/operation.rs:98
// 2^-i mod p, i = 0..32 (used to turn the O(32^2) bit recompositions of misc.rs into running sums)
ZKM_DEF_CONST(CPU_INV2, 33, {1ULL, 9223372034707292161ULL, 13835058052060938241ULL, 16140901060737761281ULL, 17293822565076172801ULL, 17870283317245378561ULL, 18158513693329981441ULL, 18302628881372282881ULL, 18374686475393433601ULL, 18410715272404008961ULL, 18428729670909296641ULL, 18437736870161940481ULL, 18442240469788262401ULL, 18444492269601423361ULL, 18445618169508003841ULL, 18446181119461294081ULL, 18446462594437939201ULL, 18446603331926261761ULL, 18446673700670423041ULL, 18446708885042503681ULL, 18446726477228544001ULL, 18446735273321564161ULL, 18446739671368074241ULL, 18446741870391329281ULL, 18446742969902956801ULL, 18446743519658770561ULL, 18446743794536677441ULL, 18446743931975630881ULL, 18446744000695107601ULL, 18446744035054845961ULL, 18446744052234715141ULL, 18446744060824649731ULL, 18446744065119617026ULL})

// util.rs:15-21 limb_from_bits_le over n consecutive columns
template <class P, class V>
ZKM_HD P bits_le(const V& lv, int start, int n) {
    // Horner from the top bit: sum_i bit_i 2^i with two additions per bit instead of a field multiplication
    // (the same field element; only the evaluation order differs)
    P s = P(0);
    ZKM_ROLLED
    for (int i = n - 1; i >= 0; i--) s = s + s + lv[start + i];
    return s;
}
template <class P>
ZKM_HD P limb_from(const P* bits, int n) {
    P s = P(0);
    for (int i = n - 1; i >= 0; i--) s = s + s + bits[i];
    return s;
}

// ---- bootstrap_kernel.rs:308-351
template <class P, class V, class YC>
ZKM_HD void eval_bootstrap_kernel(const V& lv, const V& nv, YC& yc) {
    const P local_is_bootstrap = lv[IS_BOOTSTRAP_KERNEL], next_is_bootstrap = nv[IS_BOOTSTRAP_KERNEL];
    yc.constraint_first_row(local_is_bootstrap - P(1));
    yc.constraint_last_row(local_is_bootstrap);
    const P delta = next_is_bootstrap - local_is_bootstrap;
    yc.constraint_transition(delta * (delta + P(1)));
    ZKM_ROLLED
    for (int c = 0; c < NUM_GP_CHANNELS; c++) {
        P filter = local_is_bootstrap * lv[ch(c, CH_USED)];
        yc.constraint(filter * lv[ch(c, CH_ADDR_CONTEXT)]);
        yc.constraint(filter * (lv[ch(c, CH_ADDR_SEGMENT)] - P(SEG_CODE)));
    }
    for (int c = 0; c < NUM_GP_CHANNELS; c++) yc.constraint_transition(delta * lv[ch(c, CH_USED)]);
}

// ---- decode.rs:66-95
template <class P, class V, class YC>
ZKM_HD void eval_decode(const V& lv, YC& yc) {
    const int OPCODES[8] = {OP_EQ_ISZERO, OP_KECCAK_GENERAL, OP_JUMPS, OP_BRANCH, OP_PC, OP_GET_CONTEXT, OP_SET_CONTEXT, OP_EXIT_KERNEL};
    const int COMBINED[7] = {OP_LOGIC_OP, OP_BINARY_OP, OP_BINARY_IMM_OP, OP_SHIFT, OP_SHIFT_IMM, OP_M_OP_LOAD, OP_M_OP_STORE};
    const P kernel_mode = lv[IS_KERNEL_MODE];
    yc.constraint(kernel_mode * (kernel_mode - P(1)));
    for (int i = 0; i < 6; i++) { P bit = lv[OPCODE_BITS + i]; yc.constraint(bit * (bit - P(1))); }
    for (int i = 0; i < 8; i++) { P flag = lv[OPCODES[i]]; yc.constraint(flag * (flag - P(1))); }
    for (int i = 0; i < 7; i++) { P flag = lv[COMBINED[i]]; yc.constraint(flag * (flag - P(1))); }
    P flag_sum = P(0);
    for (int i = 0; i < 8; i++) flag_sum = flag_sum + lv[OPCODES[i]];
    for (int i = 0; i < 7; i++) flag_sum = flag_sum + lv[COMBINED[i]];
    yc.constraint(flag_sum * (flag_sum - P(1)));
}

// offset = sign-extended (rd_bits[4]) 16-bit immediate << 2, as used by jumpdirect and branch
template <class P, class V>
ZKM_HD P branch_offset_bits(const V& lv) {
    P b[32];
    b[0] = P(0); b[1] = P(0);
    for (int i = 0; i < 6; i++) b[2 + i] = lv[FUNC_BITS + i];
    for (int i = 0; i < 5; i++) b[8 + i] = lv[SHAMT_BITS + i];
    for (int i = 0; i < 5; i++) b[13 + i] = lv[RD_BITS + i];
    for (int i = 18; i < 32; i++) b[i] = lv[RD_BITS + 4];
    return limb_from<P>(b, 32);
}

// ---- jumps.rs:17-240
template <class P, class V, class YC>
ZKM_HD void eval_jump_jumpi(const V& lv, const V& nv, YC& yc) {
    const P is_jump = lv[OP_JUMPS], is_jumpi = lv[OP_JUMPI], is_jumpdirect = lv[OP_JUMPDIRECT];
    const P is_link = is_jump * lv[FUNC_BITS + 0];
    const P is_linki = is_jumpi * lv[OPCODE_BITS + 0];
    {
        P reg_dst = lv[ch(0, CH_VALUE)];
        yc.constraint(is_jump * (nv[NEXT_PROGRAM_COUNTER] - reg_dst));
    }
    {
        P jump_reg = lv[ch(0, CH_ADDR_VIRTUAL)];
        P jump_dst = bits_le<P>(lv, RS_BITS, 5);
        yc.constraint(is_jump * (jump_dst - jump_reg));
    }
    {
        P b[28];
        b[0] = P(0); b[1] = P(0);
        for (int i = 0; i < 6; i++) b[2 + i] = lv[FUNC_BITS + i];
        for (int i = 0; i < 5; i++) b[8 + i] = lv[SHAMT_BITS + i];
        for (int i = 0; i < 5; i++) b[13 + i] = lv[RD_BITS + i];
        for (int i = 0; i < 5; i++) b[18 + i] = lv[RT_BITS + i];
        for (int i = 0; i < 5; i++) b[23 + i] = lv[RS_BITS + i];
        P imm_dst = limb_from<P>(b, 28);
        P pc_remain = lv[ch(2, CH_VALUE)];
        P jump_dest = pc_remain + imm_dst;
        yc.constraint(is_jumpi * (nv[NEXT_PROGRAM_COUNTER] - jump_dest));
    }
    {
        P aux = lv[ch(2, CH_VALUE)];
        const P overflow = P((uint64_t)1 << 32);
        P offset_dst = branch_offset_bits<P>(lv);
        yc.constraint(is_jumpdirect * (aux - offset_dst));
        P jump_dst = lv[PROGRAM_COUNTER] + P(4) + aux;
        yc.constraint(is_jumpdirect * (nv[NEXT_PROGRAM_COUNTER] - jump_dst) * (nv[NEXT_PROGRAM_COUNTER] + overflow - jump_dst));
    }
    {
        P link_dest = lv[ch(1, CH_VALUE)];
        yc.constraint((is_link + is_linki + is_jumpdirect) * (lv[PROGRAM_COUNTER] + P(8) - link_dest));
    }
    const P link_reg = lv[ch(1, CH_ADDR_VIRTUAL)];
    {
        P link_dst = bits_le<P>(lv, RD_BITS, 5);
        yc.constraint(is_link * (link_reg - link_dst));
    }
    yc.constraint((is_linki + is_jumpdirect) * (link_reg - P(31)));
}

// ---- jumps.rs:243-672
template <class P, class V, class YC>
ZKM_HD void eval_branch(const V& lv, const V& nv, YC& yc) {
    const P filter = lv[OP_BRANCH];
    const P is_eq = lv[BR_IS_EQ], is_ne = lv[BR_IS_NE], is_le = lv[BR_IS_LE], is_gt = lv[BR_IS_GT], is_ge = lv[BR_IS_GE], is_lt = lv[BR_IS_LT];
    const P should_jump = lv[BR_SHOULD_JUMP], b_lt = lv[BR_LT], b_gt = lv[BR_GT], b_eq = lv[BR_EQ];
    const P norm_filter = is_eq + is_ne + is_le + is_gt;
    const P special_filter = is_ge + is_lt;
    const P src1 = lv[ch(0, CH_VALUE)], src2 = lv[ch(1, CH_VALUE)], aux1 = lv[ch(2, CH_VALUE)], aux2 = lv[ch(3, CH_VALUE)],
            aux3 = lv[ch(4, CH_VALUE)], aux4 = lv[ch(5, CH_VALUE)];
    const P overflow = P((uint64_t)1 << 32), overflow_inv = P(GOLDILOCKS_INVERSE_2EXP32);
    const P one = P(1);
    yc.constraint(should_jump * (one - should_jump));
    yc.constraint(should_jump * (one - filter));
    yc.constraint(filter * (one - (norm_filter + special_filter)));
    yc.constraint(filter * (one - (b_lt + b_gt + b_eq)));
    {
        P offset_dst = branch_offset_bits<P>(lv);
        yc.constraint(filter * (aux4 - offset_dst));
        P branch_dst = lv[PROGRAM_COUNTER] + P(4) + aux4;
        yc.constraint(should_jump * (nv[NEXT_PROGRAM_COUNTER] - branch_dst) * (nv[NEXT_PROGRAM_COUNTER] + overflow - branch_dst));
        P next_inst = lv[PROGRAM_COUNTER] + P(8);
        yc.constraint(filter * (one - should_jump) * (nv[NEXT_PROGRAM_COUNTER] - next_inst));
    }
    {
        yc.constraint(filter * (aux1 + src2 - src1) * (aux1 + src2 - src1 - overflow));
        yc.constraint(filter * (aux2 + src1 - src2) * (aux2 + src1 - src2 - overflow));
        yc.constraint(filter * aux1 * ((aux1 + aux2) - overflow));
        yc.constraint(filter * aux3 * (one - aux3));
    }
    {
        P rs_reg = lv[ch(0, CH_ADDR_VIRTUAL)];
        P rs_src = bits_le<P>(lv, RS_BITS, 5);
        yc.constraint(filter * (rs_reg - rs_src));
    }
    {
        P rt_reg = lv[ch(1, CH_ADDR_VIRTUAL)];
        P rt_src = bits_le<P>(lv, RT_BITS, 5);
        yc.constraint(norm_filter * (rt_reg - rt_src));
        yc.constraint(special_filter * rt_reg * (one - rt_reg));
    }
    {
        P constr_a = src2 + aux1 - src1;
        yc.constraint(filter * constr_a * (overflow - constr_a));
        P lt0 = constr_a * overflow_inv;
        yc.constraint(b_lt * (one - lt0));
        P constr_b = src1 + aux2 - src2;
        yc.constraint(filter * constr_b * (overflow - constr_b));
        P gt0 = constr_b * overflow_inv;
        yc.constraint(b_gt * (one - gt0));
        P ne = lt0 + gt0;
        yc.constraint(b_eq * ne);
        P lt = b_lt * (one - aux3) + (one - b_lt) * aux3;
        P gt = b_gt * (one - aux3) + (one - b_gt) * aux3;
        yc.constraint(is_eq * (one - filter));
        yc.constraint(is_eq * (should_jump - (one - ne)));
        yc.constraint(is_ne * (one - filter));
        yc.constraint(is_ne * (should_jump - ne));
        yc.constraint(is_le * (one - filter));
        yc.constraint(is_le * (should_jump - (one - gt)));
        yc.constraint(is_ge * (one - filter));
        yc.constraint(is_ge * (should_jump - (one - lt)));
        yc.constraint(is_gt * (one - filter));
        yc.constraint(is_gt * (should_jump - gt));
        yc.constraint(is_lt * (one - filter));
        yc.constraint(is_lt * (should_jump - lt));
    }
}

// ---- membus.rs:34-46
template <class P, class V, class YC>
ZKM_HD void eval_membus(const V& lv, YC& yc) {
    yc.constraint(lv[CODE_CONTEXT] - (P(1) - lv[IS_KERNEL_MODE]) * lv[CONTEXT]);
    for (int c = 0; c < NUM_GP_CHANNELS; c++) { P used = lv[ch(c, CH_USED)]; yc.constraint(used * (used - P(1))); }
}

// ---- memio.rs helpers
// 32-bit word assembled from up to 4 (source column start, count) runs, low bits first; remaining bits zero
struct BitRun { int start, count; };
template <class P, class V>
ZKM_HD P word_from_runs(const V& lv, const BitRun* runs, int nruns) {
    P s = P(0);                                   // Horner from the top bit of the last run down to bit 0
    for (int r = nruns - 1; r >= 0; r--) {
        ZKM_ROLLED
        for (int i = runs[r].count - 1; i >= 0; i--) s = s + s + lv[runs[r].start + i];
    }
    return s;
}
// bits[start .. start+n) sign-extended to 32 bits (memio.rs:59-67 sign_extend::<_, N>)
template <class P, class V>
ZKM_HD P sign_extended(const V& lv, int start, int n) {
    // low n bits by Horner, plus the top bit replicated over bits n..31: top * (2^32 - 2^n)
    P s = P(0);
    ZKM_ROLLED
    for (int i = n - 1; i >= 0; i--) s = s + s + lv[start + i];
    return s + lv[start + n - 1] * P(((uint64_t)1 << 32) - ((uint64_t)1 << n));
}
// memio.rs:24-32 load_offset: 16-bit immediate (func, shamt, rd bits) sign-extended
template <class P, class V>
ZKM_HD P load_offset(const V& lv) {
    P b[32];
    for (int i = 0; i < 6; i++) b[i] = lv[FUNC_BITS + i];
    for (int i = 0; i < 5; i++) b[6 + i] = lv[SHAMT_BITS + i];
    for (int i = 0; i < 5; i++) b[11 + i] = lv[RD_BITS + i];
    for (int i = 16; i < 32; i++) b[i] = b[15];
    return limb_from<P>(b, 32);
}
// memio.rs:81-95
template <class P, class YC>
ZKM_HD void enforce_half_word(YC& yc, P op, P rs1, P mem, P mem_val_1, P mem_val_0) {
    P a = (rs1 - P(1)) * (mem - mem_val_0);
    P b = rs1 * (mem - mem_val_1);
    yc.constraint(op * (a + b));
}
// memio.rs:117-141
template <class P, class V, class YC>
ZKM_HD void enforce_byte(YC& yc, const V& lv, P op, P rs0, P rs1, P mem, P v00, P v10, P v01, P v11) {
    P prod = rs0 * rs1;
    P aux = lv[G_IO_AUX_RS0_MUL_RS1];
    yc.constraint(op * (prod - aux));
    P sum = (mem - v00) * (aux - rs1 - rs0 + P(1)) + (mem - v10) * (aux - rs0) + (mem - v01) * (aux - rs1) + (mem - v11) * aux;
    yc.constraint(sum * op);
}

template <class P, class V, class YC>
ZKM_HD void eval_memio_common(const V& lv, YC& yc, P filter, P& mem, P& rs0, P& rs1) {
    const P aux_filter = lv[MEMIO_AUX_FILTER];
    yc.constraint(filter * (P(1) - aux_filter));
    yc.constraint(filter * (lv[ch(0, CH_ADDR_SEGMENT)] - P(SEG_REGISTER_FILE)));
    yc.constraint(filter * (lv[ch(1, CH_ADDR_SEGMENT)] - P(SEG_REGISTER_FILE)));
    const P rs = lv[ch(0, CH_VALUE)], rt = lv[ch(1, CH_VALUE)];
    mem = lv[ch(3, CH_VALUE)];
    rs0 = lv[G_IO_RS_LE]; rs1 = lv[G_IO_RS_LE + 1];
    P offset = load_offset<P>(lv);
    P virt_raw = rs + offset;
    P rs_from_bits = bits_le<P>(lv, G_IO_RS_LE, 32);
    const P power32 = P((uint64_t)1 << 32);
    yc.constraint(aux_filter * (rs_from_bits - virt_raw) * (rs_from_bits + power32 - virt_raw));
    P rt_from_bits = bits_le<P>(lv, G_IO_RT_LE, 32);
    yc.constraint(filter * (rt_from_bits - rt));
    P virt = rs_from_bits - rs0 - rs1 * P(2);            // rs_limbs with bits 0 and 1 cleared
    P mem_virt = lv[ch(2, CH_ADDR_VIRTUAL)];
    yc.constraint(filter * (virt - mem_virt));
}

// ---- memio.rs:175-735 eval_packed_load
template <class P, class V, class YC>
ZKM_HD void eval_load(const V& lv, YC& yc) {
    const P filter = lv[OP_M_OP_LOAD] * lv[OPCODE_BITS + 5];
    P mem, rs0, rs1;
    eval_memio_common<P, V, YC>(lv, yc, filter, mem, rs0, rs1);
    const int RT = G_IO_RT_LE, MEM = G_IO_MEM_LE;
    {   // LH: sign-extended half words
        P mem_val_1 = sign_extended<P>(lv, MEM, 16), mem_val_0 = sign_extended<P>(lv, MEM + 16, 16);
        enforce_half_word<P, YC>(yc, lv[MEMIO_IS_LH], rs1, mem, mem_val_1, mem_val_0);
    }
    {   // LWL
        const BitRun r00[1] = {{MEM, 32}}, r10[2] = {{RT, 8}, {MEM, 24}}, r01[2] = {{RT, 16}, {MEM, 16}}, r11[2] = {{RT, 24}, {MEM, 8}};
        enforce_byte<P, V, YC>(yc, lv, lv[MEMIO_IS_LWL], rs0, rs1, mem, word_from_runs<P>(lv, r00, 1), word_from_runs<P>(lv, r10, 2),
                               word_from_runs<P>(lv, r01, 2), word_from_runs<P>(lv, r11, 2));
    }
    yc.constraint(lv[MEMIO_IS_LW] * (mem - bits_le<P>(lv, MEM, 32)));
    {   // LBU
        enforce_byte<P, V, YC>(yc, lv, lv[MEMIO_IS_LBU], rs0, rs1, mem, bits_le<P>(lv, MEM + 24, 8), bits_le<P>(lv, MEM + 16, 8),
                               bits_le<P>(lv, MEM + 8, 8), bits_le<P>(lv, MEM, 8));
    }
    {   // LHU
        P mem_val_0 = bits_le<P>(lv, MEM + 16, 16), mem_val_1 = bits_le<P>(lv, MEM, 16);
        enforce_half_word<P, YC>(yc, lv[MEMIO_IS_LHU], rs1, mem, mem_val_1, mem_val_0);
    }
    yc.checkpoint();
    {   // LWR
        const BitRun r00[2] = {{MEM + 24, 8}, {RT + 8, 24}}, r10[2] = {{MEM + 16, 16}, {RT + 16, 16}}, r01[2] = {{MEM + 8, 24}, {RT + 24, 8}},
                     r11[1] = {{MEM, 32}};
        enforce_byte<P, V, YC>(yc, lv, lv[MEMIO_IS_LWR], rs0, rs1, mem, word_from_runs<P>(lv, r00, 2), word_from_runs<P>(lv, r10, 2),
                               word_from_runs<P>(lv, r01, 2), word_from_runs<P>(lv, r11, 1));
    }
    yc.constraint(lv[MEMIO_IS_LL] * (mem - bits_le<P>(lv, MEM, 32)));
    yc.checkpoint();
    {   // LB: sign-extended bytes
        enforce_byte<P, V, YC>(yc, lv, lv[MEMIO_IS_LB], rs0, rs1, mem, sign_extended<P>(lv, MEM + 24, 8), sign_extended<P>(lv, MEM + 16, 8),
                               sign_extended<P>(lv, MEM + 8, 8), sign_extended<P>(lv, MEM, 8));
    }
    for (int c = 6; c < NUM_GP_CHANNELS - 1; c++) yc.constraint(filter * lv[ch(c, CH_USED)]);
}

// ---- memio.rs:738-1214 eval_packed_store
template <class P, class V, class YC>
ZKM_HD void eval_store(const V& lv, YC& yc) {
    const P filter = lv[OP_M_OP_STORE] * lv[OPCODE_BITS + 5];
    P mem, rs0, rs1;
    eval_memio_common<P, V, YC>(lv, yc, filter, mem, rs0, rs1);
    const int RT = G_IO_RT_LE, MEM = G_IO_MEM_LE;
    {   // SB
        const BitRun r00[2] = {{MEM, 24}, {RT, 8}}, r10[3] = {{MEM, 16}, {RT, 8}, {MEM + 24, 8}}, r01[3] = {{MEM, 8}, {RT, 8}, {MEM + 16, 16}},
                     r11[2] = {{RT, 8}, {MEM + 8, 24}};
        enforce_byte<P, V, YC>(yc, lv, lv[MEMIO_IS_SB], rs0, rs1, mem, word_from_runs<P>(lv, r00, 2), word_from_runs<P>(lv, r10, 3),
                               word_from_runs<P>(lv, r01, 3), word_from_runs<P>(lv, r11, 2));
    }
    {   // SH
        const BitRun r0[2] = {{MEM, 16}, {RT, 16}}, r1[2] = {{RT, 16}, {MEM + 16, 16}};
        enforce_half_word<P, YC>(yc, lv[MEMIO_IS_SH], rs1, mem, word_from_runs<P>(lv, r1, 2), word_from_runs<P>(lv, r0, 2));
    }
    yc.checkpoint();
    {   // SWL
        const BitRun r00[1] = {{RT, 32}}, r10[2] = {{RT + 8, 24}, {MEM + 24, 8}}, r01[2] = {{RT + 16, 16}, {MEM + 16, 16}},
                     r11[2] = {{RT + 24, 8}, {MEM + 8, 24}};
        enforce_byte<P, V, YC>(yc, lv, lv[MEMIO_IS_SWL], rs0, rs1, mem, word_from_runs<P>(lv, r00, 1), word_from_runs<P>(lv, r10, 2),
                               word_from_runs<P>(lv, r01, 2), word_from_runs<P>(lv, r11, 2));
    }
    yc.constraint(lv[MEMIO_IS_SW] * (mem - bits_le<P>(lv, RT, 32)));
    yc.checkpoint();
    {   // SWR
        const BitRun r00[2] = {{MEM, 24}, {RT, 8}}, r10[2] = {{MEM, 16}, {RT, 16}}, r01[2] = {{MEM, 8}, {RT, 24}}, r11[1] = {{RT, 32}};
        enforce_byte<P, V, YC>(yc, lv, lv[MEMIO_IS_SWR], rs0, rs1, mem, word_from_runs<P>(lv, r00, 2), word_from_runs<P>(lv, r10, 2),
                               word_from_runs<P>(lv, r01, 2), word_from_runs<P>(lv, r11, 1));
    }
    yc.constraint(lv[MEMIO_IS_SC] * (mem - bits_le<P>(lv, RT, 32)));
    yc.constraint(lv[MEMIO_IS_SDC1] * mem);
    for (int c = 6; c < NUM_GP_CHANNELS - 1; c++) yc.constraint(filter * lv[ch(c, CH_USED)]);
}

// ---- shift.rs:11-76
template <class P, class V, class YC>
ZKM_HD void eval_shift(const V& lv, YC& yc) {
    const P seg = P(SEG_SHIFT_TABLE);
    {   // variable
        const P is_shift = lv[OP_SHIFT];
        const P high_limbs_are_zero = lv[ch(3, CH_USED)];
        yc.constraint(is_shift * high_limbs_are_zero * (lv[ch(3, CH_IS_READ)] - P(1)));
        yc.constraint(is_shift * lv[ch(3, CH_ADDR_CONTEXT)]);
        yc.constraint(is_shift * (lv[ch(3, CH_ADDR_SEGMENT)] - seg));
        yc.constraint(is_shift * (lv[ch(3, CH_ADDR_VIRTUAL)] - lv[ch(0, CH_VALUE)]));
    }
    {   // immediate
        const P is_shift = lv[OP_SHIFT_IMM];
        P displacement = bits_le<P>(lv, SHAMT_BITS, 5);
        const P high_limbs_are_zero = lv[ch(3, CH_USED)];
        yc.constraint(is_shift * high_limbs_are_zero * (lv[ch(3, CH_IS_READ)] - P(1)));
        yc.constraint(is_shift * lv[ch(3, CH_ADDR_CONTEXT)]);
        yc.constraint(is_shift * (lv[ch(3, CH_ADDR_SEGMENT)] - seg));
        yc.constraint(is_shift * (lv[ch(3, CH_ADDR_VIRTUAL)] - displacement));
    }
}

// ---- count.rs:10-70
template <class P, class V, class YC>
ZKM_HD void eval_count(const V& lv, YC& yc) {
    const P filter_clz = lv[OP_CLZ_OP], filter_clo = lv[OP_CLO_OP];
    const P filter = filter_clo + filter_clz;
    yc.constraint(filter * (bits_le<P>(lv, OPCODE_BITS, 6) - P(0b011100)));
    P func = bits_le<P>(lv, FUNC_BITS, 6);
    yc.constraint(filter_clz * (func - P(0b100000)));
    yc.constraint(filter_clo * (func - P(0b100001)));
    yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RS_BITS, 5)));
    yc.constraint(filter * (lv[ch(1, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RD_BITS, 5)));
    const P rs = lv[ch(0, CH_VALUE)];
    const int BITS = G_IO_RS_LE;
    ZKM_ROLLED
    for (int i = 0; i < 32; i++) { P bit = lv[BITS + i]; yc.constraint(filter * bit * (P(1) - bit)); }
    P sum = bits_le<P>(lv, BITS, 32);
    yc.constraint(filter_clz * (rs - sum));
    yc.constraint(filter_clo * (P(0xffffffffull) - rs - sum));
    const P rd = lv[ch(1, CH_VALUE)];
    int k = 0;                                             // walks rt_le (is_eq) and mem_le (inv) together
    yc.constraint(filter * lv[BITS + 31] * rd);
    P partial = lv[BITS + 31];                             // sum_{k >= i} bit_k 2^(k-i), updated as bit_i + 2 * previous
    ZKM_ROLLED
    for (int i = 30; i >= 0; i--) {
        partial = lv[BITS + i] + partial + partial;
        P is_eq = lv[G_IO_RT_LE + k], inv = lv[G_IO_MEM_LE + k];
        k++;
        P diff = partial - P(1);
        yc.constraint(filter * diff * is_eq);
        yc.constraint(filter * (diff * inv + is_eq - P(1)));
        yc.constraint(filter * is_eq * (rd - P((uint64_t)(31 - i))));
        if (i == 0) {
            P is_eq2 = lv[G_IO_RT_LE + k], inv2 = lv[G_IO_MEM_LE + k];
            k++;
            yc.constraint(filter * partial * is_eq2);
            yc.constraint(filter * (partial * inv2 + is_eq2 - P(1)));
            yc.constraint(filter * is_eq2 * (rd - P(32)));
        }
    }
}

// ---- syscall.rs:12-222
template <class P, class V, class YC>
ZKM_HD void eval_syscall(const V& lv, YC& yc) {
    const P filter = lv[OP_SYSCALL];
    const P a0 = lv[ch(1, CH_VALUE)], a1 = lv[ch(2, CH_VALUE)], a2 = lv[ch(3, CH_VALUE)];
    const P v0 = P(0), v1 = P(0);
    const P result_v0 = lv[ch(4, CH_VALUE)], result_v1 = lv[ch(5, CH_VALUE)];
    auto cond = [&](int i) { return lv[G_SYSCALL_COND + i]; };
    auto sysnum = [&](int i) { return lv[G_SYSCALL_SYSNUM + i]; };
    auto sa0 = [&](int i) { return lv[G_SYSCALL_A0 + i]; };
    const P is_sysmap = sysnum(1);
    const P is_sz_mid_not_zero = lv[G_SYSCALL_A1];
    const P is_sz_mid_zero = sysnum(10);
    const P sz = a1;
    const P sz_in_sz_mid_not_zero = sysnum(9);
    const P is_a0_zero = sa0(0), is_a0_not_zero = sa0(2);
    const P heap_in_a0_zero = lv[ch(6, CH_VALUE)], result_heap = lv[ch(7, CH_VALUE)];
    const P is_sysmap_a0_zero = cond(0), is_sysmap_a0_zero_sz_nz = cond(1), is_sysmap_a0_zero_sz_zero = cond(2), is_sysmap_a0_nz = cond(3),
            is_sysread_a0_not_stdin = cond(4), is_sysread_a0_stdin = cond(5), is_syswrite_a0_not_stdout_err = cond(6),
            is_syswrite_a0_stdout_or_err = cond(7), is_sysfcntl_a0_stdin = cond(8), is_sysfcntl_a0_stdout_or_err = cond(9);
    const P v0_in_a0_zero = heap_in_a0_zero;
    const P heap_nz = heap_in_a0_zero + sz_in_sz_mid_not_zero, heap_z = heap_in_a0_zero + sz;
    yc.constraint(filter * (is_sysmap_a0_zero - is_sysmap * is_a0_zero));
    yc.constraint(filter * (is_sysmap_a0_zero_sz_nz - is_sysmap_a0_zero * is_sz_mid_not_zero));
    yc.constraint(filter * is_sysmap_a0_zero_sz_nz * (heap_nz - result_heap));
    yc.constraint(filter * (is_sysmap_a0_zero_sz_zero - is_sysmap_a0_zero * is_sz_mid_zero));
    yc.constraint(filter * is_sysmap_a0_zero_sz_zero * (heap_z - result_heap));
    yc.constraint(filter * is_sysmap_a0_zero * (v0_in_a0_zero - result_v0));
    yc.constraint(filter * (is_sysmap_a0_nz - is_sysmap * is_a0_not_zero));
    yc.constraint(filter * is_sysmap_a0_nz * (a0 - result_v0));
    const P is_sysbrk = sysnum(2), is_sysbrk_gt = cond(10), is_sysbrk_le = cond(11);
    const P initial_brk = lv[ch(6, CH_VALUE)];
    yc.constraint(filter * is_sysbrk * (P(1) - (is_sysbrk_gt + is_sysbrk_le)));
    yc.constraint(filter * is_sysbrk_gt * (a0 - result_v0));
    yc.constraint(filter * is_sysbrk_le * (initial_brk - result_v0));
    yc.constraint(filter * is_sysbrk * (v1 - result_v1));
    const P is_sysclone = sysnum(3);
    yc.constraint(filter * is_sysclone * (P(1) - result_v0));
    yc.constraint(filter * is_sysclone * (v1 - result_v1));
    const P is_sysread = sysnum(5);
    const P a0_is_fd_stdin = sa0(0), a0_is_not_fd_stdin = sa0(2);
    const P ffff = P(0xFFFFFFFFull), ebadf = P(MIPSEBADF);
    yc.constraint(filter * (is_sysread_a0_not_stdin - is_sysread * a0_is_not_fd_stdin));
    yc.constraint(filter * is_sysread_a0_not_stdin * (ffff - result_v0));
    yc.constraint(filter * is_sysread_a0_not_stdin * (ebadf - result_v1));
    yc.constraint(filter * (is_sysread_a0_stdin - is_sysread * a0_is_fd_stdin));
    yc.constraint(filter * is_sysread_a0_stdin * (v0 - result_v0));
    yc.constraint(filter * is_sysread_a0_stdin * (v1 - result_v1));
    const P is_syswrite = sysnum(6);
    const P a0_is_fd_stdout_or_fd_stderr = sa0(1), a0_is_not_fd_stdout_and_fd_stderr = sa0(2);
    yc.constraint(filter * (is_syswrite_a0_not_stdout_err - is_syswrite * a0_is_not_fd_stdout_and_fd_stderr));
    yc.constraint(filter * is_syswrite_a0_not_stdout_err * (ffff - result_v0));
    yc.constraint(filter * is_syswrite_a0_not_stdout_err * (ebadf - result_v1));
    yc.constraint(filter * (is_syswrite_a0_stdout_or_err - is_syswrite * a0_is_fd_stdout_or_fd_stderr));
    yc.constraint(filter * is_syswrite_a0_stdout_or_err * (a2 - result_v0));
    yc.constraint(filter * is_syswrite_a0_stdout_or_err * (v1 - result_v1));
    const P is_sysfcntl = sysnum(7);
    const P a0_is_else = sa0(2);
    yc.constraint(filter * (is_sysfcntl_a0_stdin - is_sysfcntl * a0_is_fd_stdin));
    yc.constraint(filter * is_sysfcntl_a0_stdin * (P(0) - result_v0));
    yc.constraint(filter * is_sysfcntl_a0_stdin * (v1 - result_v1));
    yc.constraint(filter * (is_sysfcntl_a0_stdout_or_err - is_sysfcntl * a0_is_fd_stdout_or_fd_stderr));
    yc.constraint(filter * is_sysfcntl_a0_stdout_or_err * (P(1) - result_v0));
    yc.constraint(filter * is_sysfcntl_a0_stdout_or_err * (v1 - result_v1));
    yc.constraint(filter * (is_sysfcntl - is_sysfcntl_a0_stdin - is_sysfcntl_a0_stdout_or_err - is_sysfcntl * a0_is_else));
    yc.constraint(filter * (is_sysfcntl - is_sysfcntl_a0_stdin - is_sysfcntl_a0_stdout_or_err) * (ffff - result_v0));
    yc.constraint(filter * (is_sysfcntl - is_sysfcntl_a0_stdin - is_sysfcntl_a0_stdout_or_err) * (ebadf - result_v1));
    const P is_syssetthreadarea = sysnum(8);
    const P threadarea = lv[ch(6, CH_VALUE)];
    yc.constraint(filter * is_syssetthreadarea * (a0 - threadarea));
}

// ---- bits.rs:9-62
template <class P, class V, class YC>
ZKM_HD void eval_bits(const V& lv, YC& yc) {
    const P filter_seh = lv[OP_SIGNEXT16], filter_seb = lv[OP_SIGNEXT8], filter_wsbh = lv[OP_SWAPHALF];
    const P filter = filter_seh + filter_seb + filter_wsbh;
    yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RT_BITS, 5)));
    yc.constraint(filter * (lv[ch(1, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RD_BITS, 5)));
    const P rt = lv[ch(0, CH_VALUE)];
    const int B = G_IO_RT_LE;
    ZKM_ROLLED
    for (int i = 0; i < 32; i++) { P bit = lv[B + i]; yc.constraint(filter * bit * (P(1) - bit)); }
    yc.constraint(filter * (rt - bits_le<P>(lv, B, 32)));
    const P rd = lv[ch(1, CH_VALUE)];
    {   // seb: bits 0..6 then bit 7 replicated
        P s = bits_le<P>(lv, B, 7) + lv[B + 7] * P(((uint64_t)1 << 32) - ((uint64_t)1 << 7));
        yc.constraint(filter_seb * (rd - s));
    }
    {   // seh
        P s = bits_le<P>(lv, B, 15) + lv[B + 15] * P(((uint64_t)1 << 32) - ((uint64_t)1 << 15));
        yc.constraint(filter_seh * (rd - s));
    }
    {   // wsbh
        const BitRun r[4] = {{B + 8, 8}, {B, 8}, {B + 24, 8}, {B + 16, 8}};
        yc.constraint(filter_wsbh * (rd - word_from_runs<P>(lv, r, 4)));
    }
}

// ---- misc.rs:10-819
template <class P, class V, class YC>
ZKM_HD void eval_misc(const V& lv, YC& yc) {
    {   // rdhwr
        const P filter = lv[OP_RDHWR];
        yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RT_BITS, 5)));
        const P rd_index = lv[G_MISC_RD_INDEX];
        yc.constraint(filter * (rd_index - bits_le<P>(lv, RD_BITS, 5)));
        const P rt_val = lv[ch(0, CH_VALUE)], local_user = lv[ch(1, CH_VALUE)];
        const P rd_eq_0 = lv[G_MISC_RD_INDEX_EQ_0], rd_eq_29 = lv[G_MISC_RD_INDEX_EQ_29];
        yc.constraint(filter * rd_eq_0 * rd_index);
        yc.constraint(filter * rd_eq_0 * (rt_val - P(1)));
        yc.constraint(filter * rd_eq_29 * (rd_index - P(29)));
        yc.constraint(filter * rd_eq_29 * (rt_val - local_user));
        yc.constraint(filter * (P(1) - rd_eq_29 - rd_eq_0) * rt_val);
    }
    yc.checkpoint();
    {   // condmov
        const P rs = lv[ch(0, CH_VALUE)], rt = lv[ch(1, CH_VALUE)], rd = lv[ch(2, CH_VALUE)], out = lv[ch(3, CH_VALUE)], mov = lv[ch(4, CH_VALUE)];
        const P is_movn = lv[OP_MOVN_OP], is_movz = lv[OP_MOVZ_OP];
        const P filter = is_movn + is_movz;
        const P is_ne = lv[G_LOGIC_DIFF_PINV] * rt;
        const P is_eq = P(1) - is_ne, no_mov = P(1) - mov;
        yc.constraint(is_movn * (mov - is_ne));
        yc.constraint(is_movz * (mov - is_eq));
        yc.constraint(filter * mov * no_mov);
        yc.constraint(filter * (out - (mov * rs + no_mov * rd)));
    }
    {   // teq
        const P filter = lv[OP_TEQ];
        yc.constraint(filter * (lv[ch(1, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RT_BITS, 5)));
        yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RS_BITS, 5)));
        const P is_ne = (lv[ch(0, CH_VALUE)] - lv[ch(1, CH_VALUE)]) * lv[G_LOGIC_DIFF_PINV];
        yc.constraint(filter * (P(1) - is_ne));
    }
    yc.checkpoint();
    {   // extract
        const P filter = lv[OP_EXT];
        yc.constraint(filter * (lv[ch(1, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RT_BITS, 5)));
        yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RS_BITS, 5)));
        const P msbd = bits_le<P>(lv, RD_BITS, 5);
        const P lsb = bits_le<P>(lv, SHAMT_BITS, 5);
        const P msb = lsb + msbd;
        const P auxm = lv[G_MISC_AUXM], auxl = lv[G_MISC_AUXL], auxs = lv[G_MISC_AUXS];
        const P rd_result = lv[ch(1, CH_VALUE)];
        yc.constraint(filter * (rd_result * auxs + auxl - auxm));
        P mpartial = P(0), lpartial = P(0);                 // sum_{j <= i} / sum_{j < i} rs_bits[j] 2^j, as running sums
        ZKM_ROLLED
        for (int i = 0; i < 32; i++) {
            lpartial = mpartial;
            mpartial = mpartial + lv[G_MISC_RS_BITS + i] * P((uint64_t)1 << i);
            P is_msb = lv[G_MISC_IS_MSB + i], is_lsb = lv[G_MISC_IS_LSB + i];
            P cur_index = P((uint64_t)i), cur_mul = P((uint64_t)1 << i);
            yc.constraint(filter * is_msb * (msb - cur_index));
            yc.constraint(filter * is_msb * (auxm - mpartial));
            yc.constraint(filter * is_lsb * (lsb - cur_index));
            yc.constraint(filter * is_lsb * (auxl - lpartial));
            yc.constraint(filter * is_lsb * (auxs - cur_mul));
        }
    }
    yc.checkpoint();
    {   // ror
        const P filter = lv[OP_ROR];
        yc.constraint(filter * (lv[ch(1, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RD_BITS, 5)));
        yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RT_BITS, 5)));
        const P sa = bits_le<P>(lv, SHAMT_BITS, 5);
        const P rd_result = lv[ch(1, CH_VALUE)];
        // rd_val_i = (bits >> i) | (low i bits << (32 - i)) = hi_i + lo_i * 2^(32-i), hi/lo as running sums
        const P full = bits_le<P>(lv, G_MISC_RS_BITS, 32);
        P lo_run = P(0);
        ZKM_ROLLED
        for (int i = 0; i < 32; i++) {
            // hi_i = sum_{k >= i} b_k 2^(k-i) = (full - lo_i) / 2^i   (exact field identity)
            P rd_val = (full - lo_run) * P(ZKM_K(CPU_INV2)[i]) + lo_run * P((uint64_t)1 << (32 - i));
            lo_run = lo_run + lv[G_MISC_RS_BITS + i] * P((uint64_t)1 << i);
            P is_sa = lv[G_MISC_IS_LSB + i];
            yc.constraint(filter * is_sa * (sa - P((uint64_t)i)));
            yc.constraint(filter * is_sa * (rd_result - rd_val));
        }
    }
    yc.checkpoint();
    {   // insert
        const P filter = lv[OP_INS];
        const P rt_src = bits_le<P>(lv, RT_BITS, 5);
        yc.constraint(filter * (lv[ch(1, CH_ADDR_VIRTUAL)] - rt_src));
        yc.constraint(filter * (lv[ch(2, CH_ADDR_VIRTUAL)] - rt_src));
        yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RS_BITS, 5)));
        const P msb = bits_le<P>(lv, RD_BITS, 5);
        const P lsb = bits_le<P>(lv, SHAMT_BITS, 5);
        const P auxm = lv[G_MISC_AUXM], auxl = lv[G_MISC_AUXL], auxs = lv[G_MISC_AUXS];
        const P rd_result = lv[ch(2, CH_VALUE)];
        yc.constraint(filter * (rd_result - auxm - auxl * auxs));
        P insert_val = P(0);
        ZKM_ROLLED
        for (int i = 0; i < 32; i++) {
            P is_msb = lv[G_MISC_IS_MSB + i], is_lsb = lv[G_MISC_IS_LSB + i];
            P cur_index = P((uint64_t)i), cur_mul = P((uint64_t)1 << i);
            yc.constraint(filter * is_lsb * (lsb - cur_index));
            yc.constraint(filter * is_lsb * (auxs - cur_mul));
            yc.constraint(filter * is_msb * (msb - lsb - cur_index));
            insert_val = insert_val + lv[G_MISC_RS_BITS + i] * P((uint64_t)1 << i);
            yc.constraint(filter * is_msb * (auxl - insert_val));
        }
    }
    yc.checkpoint();
    {   // maddu
        const P filter = lv[OP_MADDU];
        yc.constraint(filter * (lv[ch(0, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RS_BITS, 5)));
        yc.constraint(filter * (lv[ch(1, CH_ADDR_VIRTUAL)] - bits_le<P>(lv, RT_BITS, 5)));
        yc.constraint(filter * (lv[ch(2, CH_ADDR_VIRTUAL)] - P(33)));
        yc.constraint(filter * (lv[ch(4, CH_ADDR_VIRTUAL)] - P(33)));
        yc.constraint(filter * (lv[ch(3, CH_ADDR_VIRTUAL)] - P(32)));
        yc.constraint(filter * (lv[ch(5, CH_ADDR_VIRTUAL)] - P(32)));
        const P rs = lv[ch(0, CH_VALUE)], rt = lv[ch(1, CH_VALUE)], hi = lv[ch(2, CH_VALUE)], lo = lv[ch(3, CH_VALUE)],
                hi_result = lv[ch(4, CH_VALUE)], lo_result = lv[ch(5, CH_VALUE)];
        const P carry = lv[G_MISC_AUXM];
        const P scale = P((uint64_t)1 << 32);
        const P result = hi_result * scale + lo_result;
        const P mul = rs * rt;
        const P addend = hi * scale + lo;
        const P overflow = carry * scale;
        yc.constraint(filter * carry * (carry - scale));
        yc.constraint(filter * (mul + addend - overflow - result));
    }
}

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    eval_bootstrap_kernel<P, V, YC>(lv, nv, yc);
    eval_decode<P, V, YC>(lv, yc);
    yc.checkpoint();
    eval_jump_jumpi<P, V, YC>(lv, nv, yc);
    yc.checkpoint();
    eval_branch<P, V, YC>(lv, nv, yc);
    eval_membus<P, V, YC>(lv, yc);
    yc.checkpoint();
    eval_load<P, V, YC>(lv, yc);
    yc.checkpoint();
    eval_store<P, V, YC>(lv, yc);
    yc.checkpoint();
    eval_shift<P, V, YC>(lv, yc);
    eval_count<P, V, YC>(lv, yc);
    yc.checkpoint();
    eval_syscall<P, V, YC>(lv, yc);
    yc.checkpoint();
    eval_bits<P, V, YC>(lv, yc);
    yc.checkpoint();
    eval_misc<P, V, YC>(lv, yc);
    yc.checkpoint();
}

// ---- CTL selectors (cpu_stark.rs:25-244)
inline Column timestamp_col() { return Column::linear_combination({{CLOCK, (u64)NUM_CHANNELS}}); }
inline std::vector<int> opcode_func_bits() {
    std::vector<int> b = range(OPCODE_BITS, OPCODE_BITS + 6);
    for (int i = 0; i < 6; i++) b.push_back(FUNC_BITS + i);
    return b;
}
inline std::vector<Column> ctl_data_hash_sponge(int value_start, int value_count) {      // poseidon / keccak sponge
    std::vector<Column> cols = {Column::single(ch(0, CH_VALUE)), Column::single(ch(1, CH_VALUE)), Column::single(ch(2, CH_VALUE)),
                                Column::single(ch(3, CH_VALUE)), timestamp_col()};
    for (int i = 0; i < value_count; i++) cols.push_back(Column::single(value_start + i));
    return cols;
}
inline std::vector<Column> ctl_data_keccak_sponge() { return ctl_data_hash_sponge(G_KHASH_VALUE, 8); }
inline std::vector<Column> ctl_data_poseidon_sponge() { return ctl_data_hash_sponge(G_HASH_VALUE, 4); }
inline std::vector<Column> ctl_data_sha_sponge(int value_count) {                       // sha extend (1 element) / compress (8)
    std::vector<Column> cols = {Column::single(ch(0, CH_VALUE)), Column::single(ch(1, CH_VALUE)), Column::single(ch(2, CH_VALUE)),
                                timestamp_col()};
    for (int i = 0; i < value_count; i++) cols.push_back(Column::single(GENERAL + i));
    return cols;
}
inline std::vector<Column> ctl_data_sha_extend_sponge() { return ctl_data_sha_sponge(1); }
inline std::vector<Column> ctl_data_sha_compress_sponge() { return ctl_data_sha_sponge(8); }
inline Filter ctl_filter_keccak_sponge() { return Filter::new_simple(Column::single(IS_KECCAK_SPONGE)); }
inline Filter ctl_filter_poseidon_sponge() { return Filter::new_simple(Column::single(IS_POSEIDON_SPONGE)); }
inline Filter ctl_filter_sha_extend_sponge() { return Filter::new_simple(Column::single(IS_SHA_EXTEND_SPONGE)); }
inline Filter ctl_filter_sha_compress_sponge() { return Filter::new_simple(Column::single(IS_SHA_COMPRESS_SPONGE)); }
inline std::vector<Column> ctl_data_binops() {
    return {Column::single(ch(0, CH_VALUE)), Column::single(ch(1, CH_VALUE)), Column::single(ch(2, CH_VALUE))};
}
inline std::vector<Column> ctl_data_logic() {
    std::vector<Column> res = {Column::le_bits(opcode_func_bits())};
    for (auto& c : ctl_data_binops()) res.push_back(c);
    return res;
}
inline Filter ctl_filter_logic() { return Filter::new_simple(Column::single(OP_LOGIC_OP)); }
inline TableWithColumns ctl_arithmetic_base_rows(int table) {
    std::vector<Column> cols = {Column::le_bits(opcode_func_bits())};
    for (auto& c : ctl_data_binops()) cols.push_back(c);
    return TableWithColumns(table, cols, Filter::new_simple(Column::sum({OP_BINARY_OP, OP_SHIFT, OP_SHIFT_IMM})));
}
inline TableWithColumns ctl_arithmetic_imm_base_rows(int table) {
    std::vector<Column> cols = {Column::le_bits(range(OPCODE_BITS, OPCODE_BITS + 6))};
    for (auto& c : ctl_data_binops()) cols.push_back(c);
    return TableWithColumns(table, cols, Filter::new_simple(Column::single(OP_BINARY_IMM_OP)));
}
inline Column mem_time_and_channel(int channel) {
    return Column::linear_combination_with_constant({{CLOCK, (u64)NUM_CHANNELS}}, (u64)channel);
}
inline std::vector<Column> ctl_data_code_memory() {
    return {Column::constant_(1), Column::single(CODE_CONTEXT), Column::constant_(SEG_CODE), Column::single(PROGRAM_COUNTER),
            Column::le_bits(opcode_func_bits()), mem_time_and_channel(0)};
}
inline std::vector<Column> ctl_data_gp_memory(int channel) {
    std::vector<Column> cols = Column::singles({ch(channel, CH_IS_READ), ch(channel, CH_ADDR_CONTEXT), ch(channel, CH_ADDR_SEGMENT),
                                                ch(channel, CH_ADDR_VIRTUAL), ch(channel, CH_VALUE)});
    cols.push_back(mem_time_and_channel(0));
    return cols;
}
inline Filter ctl_filter_gp_memory(int channel) { return Filter::new_simple(Column::single(ch(channel, CH_USED))); }

}  // namespace cpu
}  // namespace tables
}  // namespace zkm
