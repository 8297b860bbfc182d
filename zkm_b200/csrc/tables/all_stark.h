// The 12-table AllStark System: table order = the reference's `Table` enum (prover/src/all_stark.rs:
// 97-133) and the 15 cross-table lookups of all_cross_table_lookups (:136-154) in the same order, each
// built exactly as its ctl_* function (:156-542).
#pragma once
#include "registry.h"

namespace zkm {
namespace tables {

inline System all_stark_system() {
    System s;
    for (int k = 0; k < NUM_TABLE_KINDS; k++) s.kinds.push_back(k);      // table index == kind
    auto ctl = [](std::vector<TableWithColumns> looking, TableWithColumns looked) {
        CrossTableLookup c;
        c.looking_tables = std::move(looking);
        c.looked_table = std::move(looked);
        return c;
    };
    // ctl_arithmetic (:156-164)
    s.ctls.push_back(ctl({cpu::ctl_arithmetic_base_rows(T_CPU), cpu::ctl_arithmetic_imm_base_rows(T_CPU)},
                         arithmetic::ctl_arithmetic_rows(T_ARITHMETIC)));
    // ctl_poseidon_sponge (:199-211)
    s.ctls.push_back(ctl({TableWithColumns(T_CPU, cpu::ctl_data_poseidon_sponge(), cpu::ctl_filter_poseidon_sponge())},
                         TableWithColumns(T_POSEIDON_SPONGE, poseidon_sponge::ctl_looked_data(), poseidon_sponge::ctl_looked_filter())));
    // ctl_poseidon_inputs / outputs (:169-197)
    s.ctls.push_back(ctl({TableWithColumns(T_POSEIDON_SPONGE, poseidon_sponge::ctl_looking_poseidon_inputs(), poseidon_sponge::ctl_looking_poseidon_filter())},
                         TableWithColumns(T_POSEIDON, poseidon::ctl_data_inputs(), poseidon::ctl_filter_inputs())));
    s.ctls.push_back(ctl({TableWithColumns(T_POSEIDON_SPONGE, poseidon_sponge::ctl_looking_poseidon_outputs(), poseidon_sponge::ctl_looking_poseidon_filter())},
                         TableWithColumns(T_POSEIDON, poseidon::ctl_data_outputs(), poseidon::ctl_filter_outputs())));
    // ctl_keccak_sponge (:244-256), ctl_keccak_inputs / outputs (:216-242)
    s.ctls.push_back(ctl({TableWithColumns(T_CPU, cpu::ctl_data_keccak_sponge(), cpu::ctl_filter_keccak_sponge())},
                         TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looked_data(), keccak_sponge::ctl_looked_filter())));
    s.ctls.push_back(ctl({TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_keccak_inputs(), keccak_sponge::ctl_looking_keccak_filter())},
                         TableWithColumns(T_KECCAK, keccak::ctl_data_inputs(), keccak::ctl_filter_inputs())));
    s.ctls.push_back(ctl({TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_keccak_outputs(), keccak_sponge::ctl_looking_keccak_filter())},
                         TableWithColumns(T_KECCAK, keccak::ctl_data_outputs(), keccak::ctl_filter_outputs())));
    // ctl_sha_extend_sponge (:286-298), ctl_sha_extend_inputs / outputs (:258-284)
    s.ctls.push_back(ctl({TableWithColumns(T_CPU, cpu::ctl_data_sha_extend_sponge(), cpu::ctl_filter_sha_extend_sponge())},
                         TableWithColumns(T_SHA_EXTEND_SPONGE, sha_extend_sponge::ctl_looked_data(), sha_extend_sponge::ctl_looking_sha_extend_filter())));
    s.ctls.push_back(ctl({TableWithColumns(T_SHA_EXTEND_SPONGE, sha_extend_sponge::ctl_looking_sha_extend_inputs(), sha_extend_sponge::ctl_looking_sha_extend_filter())},
                         TableWithColumns(T_SHA_EXTEND, sha_extend::ctl_data_inputs(), sha_extend::ctl_filter())));
    s.ctls.push_back(ctl({TableWithColumns(T_SHA_EXTEND_SPONGE, sha_extend_sponge::ctl_looking_sha_extend_outputs(), sha_extend_sponge::ctl_looking_sha_extend_filter())},
                         TableWithColumns(T_SHA_EXTEND, sha_extend::ctl_data_outputs(), sha_extend::ctl_filter())));
    // ctl_sha_compress_sponge (:328-340), ctl_sha_compress_inputs / outputs (:300-326)
    s.ctls.push_back(ctl({TableWithColumns(T_CPU, cpu::ctl_data_sha_compress_sponge(), cpu::ctl_filter_sha_compress_sponge())},
                         TableWithColumns(T_SHA_COMPRESS_SPONGE, sha_compress_sponge::ctl_looked_data(), sha_compress_sponge::ctl_looked_filter())));
    s.ctls.push_back(ctl({TableWithColumns(T_SHA_COMPRESS_SPONGE, sha_compress_sponge::ctl_looking_sha_compress_inputs(),
                                           sha_compress_sponge::ctl_looking_sha_compress_filter())},
                         TableWithColumns(T_SHA_COMPRESS, sha_compress::ctl_data_inputs(), sha_compress::ctl_filter_inputs())));
    s.ctls.push_back(ctl({TableWithColumns(T_SHA_COMPRESS_SPONGE, sha_compress_sponge::ctl_looking_sha_compress_outputs(),
                                           sha_compress_sponge::ctl_looking_sha_compress_filter())},
                         TableWithColumns(T_SHA_COMPRESS, sha_compress::ctl_data_outputs(), sha_compress::ctl_filter_outputs())));
    {   // ctl_logic (:340-477)
        std::vector<TableWithColumns> lookers;
        lookers.push_back(TableWithColumns(T_CPU, cpu::ctl_data_logic(), cpu::ctl_filter_logic()));
        for (int i = 0; i < keccak_sponge::num_logic_ctls(); i++)
            lookers.push_back(TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_logic(i), keccak_sponge::ctl_looking_logic_filter()));
        lookers.push_back(TableWithColumns(T_SHA_EXTEND, sha_extend::ctl_s_0_inter_looking_logic(), sha_extend::ctl_filter()));
        lookers.push_back(TableWithColumns(T_SHA_EXTEND, sha_extend::ctl_s_0_looking_logic(), sha_extend::ctl_filter()));
        lookers.push_back(TableWithColumns(T_SHA_EXTEND, sha_extend::ctl_s_1_inter_looking_logic(), sha_extend::ctl_filter()));
        lookers.push_back(TableWithColumns(T_SHA_EXTEND, sha_extend::ctl_s_1_looking_logic(), sha_extend::ctl_filter()));
        typedef std::vector<Column> (*colfn)();
        const colfn sc[12] = {sha_compress::ctl_s_1_inter_looking_logic, sha_compress::ctl_s_1_looking_logic, sha_compress::ctl_e_and_f_looking_logic,
                              sha_compress::ctl_not_e_and_g_looking_logic, sha_compress::ctl_ch_looking_logic, sha_compress::ctl_s_0_inter_looking_logic,
                              sha_compress::ctl_s_0_looking_logic, sha_compress::ctl_a_and_b_looking_logic, sha_compress::ctl_a_and_c_looking_logic,
                              sha_compress::ctl_b_and_c_looking_logic, sha_compress::ctl_maj_inter_looking_logic, sha_compress::ctl_maj_looking_logic};
        for (colfn f : sc) lookers.push_back(TableWithColumns(T_SHA_COMPRESS, f(), sha_compress::ctl_logic_filter()));
        s.ctls.push_back(ctl(lookers, TableWithColumns(T_LOGIC, logic::ctl_data(), logic::ctl_filter())));
    }
    {   // ctl_memory (:479-542): cpu gp channels, keccak sponge, poseidon sponge, sha extend sponge, sha compress sponge, sha compress
        std::vector<TableWithColumns> lookers;
        for (int c = 0; c < cpu::NUM_GP_CHANNELS; c++) lookers.push_back(TableWithColumns(T_CPU, cpu::ctl_data_gp_memory(c), cpu::ctl_filter_gp_memory(c)));
        for (int i = 0; i < keccak_sponge::KECCAK_RATE_BYTES; i++)
            lookers.push_back(TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_memory(i), keccak_sponge::ctl_looking_memory_filter(i)));
        for (int i = 0; i < poseidon_sponge::POSEIDON_RATE_BYTES; i++)
            lookers.push_back(TableWithColumns(T_POSEIDON_SPONGE, poseidon_sponge::ctl_looking_memory(i), poseidon_sponge::ctl_looking_memory_filter(i)));
        for (int i = 0; i < sha_extend_sponge::SHA_EXTEND_SPONGE_READ_BYTES; i++)
            lookers.push_back(TableWithColumns(T_SHA_EXTEND_SPONGE, sha_extend_sponge::ctl_looking_memory(i), sha_extend_sponge::ctl_looking_sha_extend_filter()));
        for (int i = 0; i < sha_compress_sponge::SHA_COMPRESS_SPONGE_READ_BYTES; i++)
            lookers.push_back(TableWithColumns(T_SHA_COMPRESS_SPONGE, sha_compress_sponge::ctl_looking_memory(i),
                                               sha_compress_sponge::ctl_looking_sha_compress_filter()));
        for (int i = 0; i < 4; i++) lookers.push_back(TableWithColumns(T_SHA_COMPRESS, sha_compress::ctl_looking_memory(i), sha_compress::ctl_logic_filter()));
        s.ctls.push_back(ctl(lookers, TableWithColumns(T_MEMORY, memory::ctl_data(), memory::ctl_filter())));
    }
    return s;
}

}  // namespace tables
}  // namespace zkm
