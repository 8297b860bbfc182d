// Memory table (13 columns): address/timestamp-sorted memory log with first-change flags and a
// logUp range check.  Column map: reference prover/src/memory/columns.rs:6-37 (VALUE_LIMBS = 1,
// memory/mod.rs); constraints: memory/memory_stark.rs:256-341 (the dummy-write constraint at
// :289-296 is commented out in the reference and therefore absent here); lookup :476-483; CTL
// selectors :29-39.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace memory {

constexpr int FILTER = 0, TIMESTAMP = 1, IS_READ = 2, ADDR_CONTEXT = 3, ADDR_SEGMENT = 4, ADDR_VIRTUAL = 5;
constexpr int VALUE_LIMBS = 1;
constexpr int VALUE_START = 6;
constexpr int CONTEXT_FIRST_CHANGE = VALUE_START + VALUE_LIMBS, SEGMENT_FIRST_CHANGE = CONTEXT_FIRST_CHANGE + 1,
              VIRTUAL_FIRST_CHANGE = SEGMENT_FIRST_CHANGE + 1, RANGE_CHECK = VIRTUAL_FIRST_CHANGE + 1,
              COUNTER = RANGE_CHECK + 1, FREQUENCIES = COUNTER + 1;
constexpr int NUM_COLUMNS = FREQUENCIES + 1;

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    const P one = P(1);
    const P timestamp = lv[TIMESTAMP], addr_context = lv[ADDR_CONTEXT], addr_segment = lv[ADDR_SEGMENT],
            addr_virtual = lv[ADDR_VIRTUAL];
    const P next_timestamp = nv[TIMESTAMP], next_is_read = nv[IS_READ], next_addr_context = nv[ADDR_CONTEXT],
            next_addr_segment = nv[ADDR_SEGMENT], next_addr_virtual = nv[ADDR_VIRTUAL];

    const P filter = lv[FILTER];
    yc.constraint(filter * (filter - one));

    const P context_first_change = lv[CONTEXT_FIRST_CHANGE], segment_first_change = lv[SEGMENT_FIRST_CHANGE],
            virtual_first_change = lv[VIRTUAL_FIRST_CHANGE];
    const P address_unchanged = one - context_first_change - segment_first_change - virtual_first_change;
    const P range_check = lv[RANGE_CHECK];
    const P not_context_first_change = one - context_first_change, not_segment_first_change = one - segment_first_change,
            not_virtual_first_change = one - virtual_first_change, not_address_unchanged = one - address_unchanged;

    yc.constraint(context_first_change * not_context_first_change);
    yc.constraint(segment_first_change * not_segment_first_change);
    yc.constraint(virtual_first_change * not_virtual_first_change);
    yc.constraint(address_unchanged * not_address_unchanged);

    yc.constraint_transition(segment_first_change * (next_addr_context - addr_context));
    yc.constraint_transition(virtual_first_change * (next_addr_context - addr_context));
    yc.constraint_transition(virtual_first_change * (next_addr_segment - addr_segment));
    yc.constraint_transition(address_unchanged * (next_addr_context - addr_context));
    yc.constraint_transition(address_unchanged * (next_addr_segment - addr_segment));
    yc.constraint_transition(address_unchanged * (next_addr_virtual - addr_virtual));

    const P computed_range_check = context_first_change * (next_addr_context - addr_context - one) +
                                   segment_first_change * (next_addr_segment - addr_segment - one) +
                                   virtual_first_change * (next_addr_virtual - addr_virtual - one) +
                                   address_unchanged * (next_timestamp - timestamp);
    yc.constraint_transition(range_check - computed_range_check);

    for (int i = 0; i < VALUE_LIMBS; i++)
        yc.constraint_transition(next_is_read * address_unchanged * (nv[VALUE_START + i] - lv[VALUE_START + i]));
}

inline std::vector<Column> ctl_data() {
    std::vector<Column> res = Column::singles({IS_READ, ADDR_CONTEXT, ADDR_SEGMENT, ADDR_VIRTUAL});
    for (int i = 0; i < VALUE_LIMBS; i++) res.push_back(Column::single(VALUE_START + i));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline Filter ctl_filter() { return Filter::new_simple(Column::single(FILTER)); }

inline std::vector<Lookup> lookups() {
    Lookup l;
    l.columns = {Column::single(RANGE_CHECK)};
    l.table_column = Column::single(COUNTER);
    l.frequencies_column = Column::single(FREQUENCIES);
    l.filter_columns = {Filter::none()};
    return {l};
}

}  // namespace memory
}  // namespace tables
}  // namespace zkm
