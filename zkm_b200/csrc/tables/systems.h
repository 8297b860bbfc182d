// The Systems the library can prove.  SYSTEM_ALL_STARK is the reference's AllStark (all_stark.rs);
// the others are small self-contained Systems (same code path, fewer tables) whose valid traces can
// be generated without the MIPS emulator, used for end-to-end prove -> verify parity tests in the
// spirit of the reference's single-table prove tests (poseidon_stark.rs:751-816, keccak_stark.rs:689-754).
#pragma once
#include "all_stark.h"

namespace zkm {
namespace tables {

enum SystemId { SYSTEM_ALL_STARK = 0, SYSTEM_LOGIC = 1, SYSTEM_MINI3 = 2, SYSTEM_POSEIDON = 3, SYSTEM_MEMORY = 4, SYSTEM_ARITH = 5, SYSTEM_KECCAK = 6, SYSTEM_POSEIDON_SPONGE = 7, SYSTEM_SHA_EXTEND = 8, SYSTEM_SHA_COMPRESS = 9, SYSTEM_CPU = 10 };

// A table looked up by itself: looking = looked (multiset equality holds trivially for any trace).
inline CrossTableLookup self_ctl(int table, std::vector<Column> cols, Filter f) {
    CrossTableLookup c;
    c.looking_tables.push_back(TableWithColumns(table, cols, f));
    c.looked_table = TableWithColumns(table, cols, f);
    return c;
}

inline System make_system(int id) {
    System s;
    switch (id) {
        case SYSTEM_ALL_STARK:
            return all_stark_system();
        case SYSTEM_LOGIC:
            s.kinds = {T_LOGIC};
            s.ctls.push_back(self_ctl(0, logic::ctl_data(), logic::ctl_filter()));
            return s;
        case SYSTEM_POSEIDON:
            s.kinds = {T_POSEIDON};
            s.ctls.push_back(self_ctl(0, poseidon::ctl_data_inputs(), poseidon::ctl_filter_inputs()));
            s.ctls.push_back(self_ctl(0, poseidon::ctl_data_outputs(), poseidon::ctl_filter_outputs()));
            return s;
        case SYSTEM_MEMORY:
            s.kinds = {T_MEMORY};
            s.ctls.push_back(self_ctl(0, memory::ctl_data(), memory::ctl_filter()));
            return s;
        case SYSTEM_ARITH: {
            // Arithmetic table alone: its 18-column range-check logUp (arithmetic_stark.rs:269-276) plus the CPU-facing CTL
            // rows (arithmetic_stark.rs:61-116) looked up by themselves.
            s.kinds = {T_ARITHMETIC};
            TableWithColumns rows = arithmetic::ctl_arithmetic_rows(0);
            s.ctls.push_back(self_ctl(0, rows.columns, rows.filter));
            return s;
        }
        case SYSTEM_KECCAK: {
            // The Keccak slice of AllStark with its real CTLs (all_stark.rs:216-256,356-386,479-542 restricted to these
            // tables): 0 = Keccak, 1 = KeccakSponge, 2 = Logic, 3 = Memory.  The sponge's absorbed blocks go through the
            // permutation table (inputs, outputs), its 34 XORs per row through Logic, its 136 byte reads through Memory;
            // the CPU-facing sponge output rows are looked up by themselves.
            s.kinds = {T_KECCAK, T_KECCAK_SPONGE, T_LOGIC, T_MEMORY};
            CrossTableLookup in, out, lg, mem;
            in.looking_tables.push_back(TableWithColumns(1, keccak_sponge::ctl_looking_keccak_inputs(), keccak_sponge::ctl_looking_keccak_filter()));
            in.looked_table = TableWithColumns(0, keccak::ctl_data_inputs(), keccak::ctl_filter_inputs());
            out.looking_tables.push_back(TableWithColumns(1, keccak_sponge::ctl_looking_keccak_outputs(), keccak_sponge::ctl_looking_keccak_filter()));
            out.looked_table = TableWithColumns(0, keccak::ctl_data_outputs(), keccak::ctl_filter_outputs());
            for (int i = 0; i < keccak_sponge::num_logic_ctls(); i++)
                lg.looking_tables.push_back(TableWithColumns(1, keccak_sponge::ctl_looking_logic(i), keccak_sponge::ctl_looking_logic_filter()));
            lg.looked_table = TableWithColumns(2, logic::ctl_data(), logic::ctl_filter());
            for (int i = 0; i < keccak_sponge::KECCAK_RATE_BYTES; i++)
                mem.looking_tables.push_back(TableWithColumns(1, keccak_sponge::ctl_looking_memory(i), keccak_sponge::ctl_looking_memory_filter(i)));
            mem.looked_table = TableWithColumns(3, memory::ctl_data(), memory::ctl_filter());
            s.ctls = {in, out, lg, mem, self_ctl(1, keccak_sponge::ctl_looked_data(), keccak_sponge::ctl_looked_filter())};
            return s;
        }
        case SYSTEM_POSEIDON_SPONGE: {
            // The Poseidon slice of AllStark (all_stark.rs:169-211,479-542): 0 = Poseidon, 1 = PoseidonSponge, 2 = Memory.
            s.kinds = {T_POSEIDON, T_POSEIDON_SPONGE, T_MEMORY};
            CrossTableLookup in, out, mem;
            in.looking_tables.push_back(TableWithColumns(1, poseidon_sponge::ctl_looking_poseidon_inputs(), poseidon_sponge::ctl_looking_poseidon_filter()));
            in.looked_table = TableWithColumns(0, poseidon::ctl_data_inputs(), poseidon::ctl_filter_inputs());
            out.looking_tables.push_back(TableWithColumns(1, poseidon_sponge::ctl_looking_poseidon_outputs(), poseidon_sponge::ctl_looking_poseidon_filter()));
            out.looked_table = TableWithColumns(0, poseidon::ctl_data_outputs(), poseidon::ctl_filter_outputs());
            for (int i = 0; i < poseidon_sponge::POSEIDON_RATE_BYTES; i++)
                mem.looking_tables.push_back(TableWithColumns(1, poseidon_sponge::ctl_looking_memory(i), poseidon_sponge::ctl_looking_memory_filter(i)));
            mem.looked_table = TableWithColumns(2, memory::ctl_data(), memory::ctl_filter());
            s.ctls = {in, out, mem, self_ctl(1, poseidon_sponge::ctl_looked_data(), poseidon_sponge::ctl_looked_filter())};
            return s;
        }
        case SYSTEM_SHA_EXTEND: {
            // The SHA-256 message-schedule slice of AllStark (all_stark.rs:258-298,356-386,479-542):
            // 0 = ShaExtend, 1 = ShaExtendSponge, 2 = Logic, 3 = Memory.
            s.kinds = {T_SHA_EXTEND, T_SHA_EXTEND_SPONGE, T_LOGIC, T_MEMORY};
            CrossTableLookup in, out, lg, mem;
            in.looking_tables.push_back(TableWithColumns(1, sha_extend_sponge::ctl_looking_sha_extend_inputs(), sha_extend_sponge::ctl_looking_sha_extend_filter()));
            in.looked_table = TableWithColumns(0, sha_extend::ctl_data_inputs(), sha_extend::ctl_filter());
            out.looking_tables.push_back(TableWithColumns(1, sha_extend_sponge::ctl_looking_sha_extend_outputs(), sha_extend_sponge::ctl_looking_sha_extend_filter()));
            out.looked_table = TableWithColumns(0, sha_extend::ctl_data_outputs(), sha_extend::ctl_filter());
            lg.looking_tables = {TableWithColumns(0, sha_extend::ctl_s_0_inter_looking_logic(), sha_extend::ctl_filter()),
                                 TableWithColumns(0, sha_extend::ctl_s_0_looking_logic(), sha_extend::ctl_filter()),
                                 TableWithColumns(0, sha_extend::ctl_s_1_inter_looking_logic(), sha_extend::ctl_filter()),
                                 TableWithColumns(0, sha_extend::ctl_s_1_looking_logic(), sha_extend::ctl_filter())};
            lg.looked_table = TableWithColumns(2, logic::ctl_data(), logic::ctl_filter());
            for (int i = 0; i < sha_extend_sponge::SHA_EXTEND_SPONGE_READ_BYTES; i++)
                mem.looking_tables.push_back(TableWithColumns(1, sha_extend_sponge::ctl_looking_memory(i), sha_extend_sponge::ctl_looking_sha_extend_filter()));
            mem.looked_table = TableWithColumns(3, memory::ctl_data(), memory::ctl_filter());
            s.ctls = {in, out, lg, mem, self_ctl(1, sha_extend_sponge::ctl_looked_data(), sha_extend_sponge::ctl_looking_sha_extend_filter())};
            return s;
        }
        case SYSTEM_SHA_COMPRESS: {
            // The SHA-256 compression slice of AllStark (all_stark.rs:300-340,356-386,479-542):
            // 0 = ShaCompress, 1 = ShaCompressSponge, 2 = Logic, 3 = Memory.
            s.kinds = {T_SHA_COMPRESS, T_SHA_COMPRESS_SPONGE, T_LOGIC, T_MEMORY};
            CrossTableLookup in, out, lg, mem;
            in.looking_tables.push_back(TableWithColumns(1, sha_compress_sponge::ctl_looking_sha_compress_inputs(), sha_compress_sponge::ctl_looking_sha_compress_filter()));
            in.looked_table = TableWithColumns(0, sha_compress::ctl_data_inputs(), sha_compress::ctl_filter_inputs());
            out.looking_tables.push_back(TableWithColumns(1, sha_compress_sponge::ctl_looking_sha_compress_outputs(), sha_compress_sponge::ctl_looking_sha_compress_filter()));
            out.looked_table = TableWithColumns(0, sha_compress::ctl_data_outputs(), sha_compress::ctl_filter_outputs());
            typedef std::vector<Column> (*colfn)();
            const colfn sc[12] = {sha_compress::ctl_s_1_inter_looking_logic, sha_compress::ctl_s_1_looking_logic, sha_compress::ctl_e_and_f_looking_logic,
                                  sha_compress::ctl_not_e_and_g_looking_logic, sha_compress::ctl_ch_looking_logic, sha_compress::ctl_s_0_inter_looking_logic,
                                  sha_compress::ctl_s_0_looking_logic, sha_compress::ctl_a_and_b_looking_logic, sha_compress::ctl_a_and_c_looking_logic,
                                  sha_compress::ctl_b_and_c_looking_logic, sha_compress::ctl_maj_inter_looking_logic, sha_compress::ctl_maj_looking_logic};
            for (colfn f : sc) lg.looking_tables.push_back(TableWithColumns(0, f(), sha_compress::ctl_logic_filter()));
            lg.looked_table = TableWithColumns(2, logic::ctl_data(), logic::ctl_filter());
            for (int i = 0; i < sha_compress_sponge::SHA_COMPRESS_SPONGE_READ_BYTES; i++)
                mem.looking_tables.push_back(TableWithColumns(1, sha_compress_sponge::ctl_looking_memory(i), sha_compress_sponge::ctl_looking_sha_compress_filter()));
            for (int i = 0; i < 4; i++) mem.looking_tables.push_back(TableWithColumns(0, sha_compress::ctl_looking_memory(i), sha_compress::ctl_logic_filter()));
            mem.looked_table = TableWithColumns(3, memory::ctl_data(), memory::ctl_filter());
            s.ctls = {in, out, lg, mem, self_ctl(1, sha_compress_sponge::ctl_looked_data(), sha_compress_sponge::ctl_looked_filter())};
            return s;
        }
        case SYSTEM_CPU: {
            // The instruction-execution slice of AllStark (all_stark.rs:156-164, 340-355, 479-487):
            // 0 = Cpu, 1 = Arithmetic, 2 = Logic, 3 = Memory with the CPU's arithmetic, logic and 9 memory-channel lookups.
            s.kinds = {T_CPU, T_ARITHMETIC, T_LOGIC, T_MEMORY};
            CrossTableLookup ar, lg, mem;
            ar.looking_tables = {cpu::ctl_arithmetic_base_rows(0), cpu::ctl_arithmetic_imm_base_rows(0)};
            ar.looked_table = arithmetic::ctl_arithmetic_rows(1);
            lg.looking_tables = {TableWithColumns(0, cpu::ctl_data_logic(), cpu::ctl_filter_logic())};
            lg.looked_table = TableWithColumns(2, logic::ctl_data(), logic::ctl_filter());
            for (int c = 0; c < cpu::NUM_GP_CHANNELS; c++) mem.looking_tables.push_back(TableWithColumns(0, cpu::ctl_data_gp_memory(c), cpu::ctl_filter_gp_memory(c)));
            mem.looked_table = TableWithColumns(3, memory::ctl_data(), memory::ctl_filter());
            s.ctls = {ar, lg, mem};
            return s;
        }
        case SYSTEM_MINI3: {
            // tables: 0 = Poseidon, 1 = Logic, 2 = Memory
            s.kinds = {T_POSEIDON, T_LOGIC, T_MEMORY};
            // Logic looked by three looking sets of the same table (-> 2 helper columns): rows with
            // IS_AND, rows with IS_OR, rows with IS_XOR + IS_NOR; their union is the looked filter.
            CrossTableLookup a;
            a.looking_tables.push_back(TableWithColumns(1, logic::ctl_data(), Filter::new_simple(Column::single(logic::IS_AND))));
            a.looking_tables.push_back(TableWithColumns(1, logic::ctl_data(), Filter::new_simple(Column::single(logic::IS_OR))));
            a.looking_tables.push_back(TableWithColumns(1, logic::ctl_data(), Filter::new_simple(Column::sum({logic::IS_XOR, logic::IS_NOR}))));
            a.looked_table = TableWithColumns(1, logic::ctl_data(), logic::ctl_filter());
            s.ctls.push_back(a);
            s.ctls.push_back(self_ctl(0, poseidon::ctl_data_inputs(), poseidon::ctl_filter_inputs()));
            s.ctls.push_back(self_ctl(0, poseidon::ctl_data_outputs(), poseidon::ctl_filter_outputs()));
            // Memory looked by two looking sets of itself (reads, writes)
            CrossTableLookup m;
            m.looking_tables.push_back(TableWithColumns(2, memory::ctl_data(), Filter::new_(
                {{Column::single(memory::FILTER), Column::single(memory::IS_READ)}}, {})));
            m.looking_tables.push_back(TableWithColumns(2, memory::ctl_data(), Filter::new_(
                {{Column::single(memory::FILTER), Column::linear_combination_with_constant({{memory::IS_READ, FP - 1}}, 1)}}, {})));
            m.looked_table = TableWithColumns(2, memory::ctl_data(), memory::ctl_filter());
            s.ctls.push_back(m);
            return s;
        }
        default:
            throw std::runtime_error("unknown or not yet available system id");
    }
}

}  // namespace tables
}  // namespace zkm
