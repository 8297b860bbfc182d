"""The ordering argument of shim/src/b200_split.rs, checked on a model (CPU, no device, no Poseidon: the argument does not depend
on the hash function, so a 32-byte blake2s stands in for hash_page).

`RefMemory` restates the page bookkeeping of the reference's emulator (emulator/src/memory.rs: get_memory :221-250, set_memory
:262-296, set_hash_trace :203-219, alloc_hash_page :378-386, set_hash_range :388-413, update_page_hash :415-436, compute_image_id
:438-471, get_input_image :524-538) and `split_reference` the body of InstrumentedState::split_segment (state.rs:1477-1530).
`split_adaptor` is what the Rust adaptor does instead: the two INIT_SP reads first, then ONE call that updates the hash pages and
computes the image id in a separate page tree (`TreeModel` = the documented behaviour of zkm_b200_splitter_split /
zkm_pagetree_t), then the touched hash pages copied back.  On random programs over several segments both must write the same
segment images, image ids and roots, and leave the same memory behind."""
import hashlib

import numpy as np
import pytest

PAGE, INIT_SP, MAX_MEMORY, ROOT_PAGE, REGISTERS_OFFSET = 4096, 0x7FFFD000, 0x80000000, 0x81020, 0x400
CONST = [bytes([0xA0 + lv]) * PAGE for lv in range(3)]           # stand-ins for CONST_HASH_PAGES[level]


def hash_page(data):
    return hashlib.blake2s(bytes(data)).digest()


class RefMemory:
    def __init__(self):
        self.pages, self.rtrace, self.wtrace = {}, {}, [{}, {}, {}]

    def alloc_hash_page(self, page_index, level):
        self.pages[page_index] = bytearray(CONST[level])
        return self.pages[page_index]

    def set_hash_trace(self, page_index, level):
        hp = ((page_index << 5) + MAX_MEMORY) >> 12
        page = self.pages.get(hp)
        if page is None:
            page = self.alloc_hash_page(hp, level)
        self.rtrace.setdefault(hp, bytes(page))
        if level < 2:
            self.set_hash_trace(hp, level + 1)

    def get_memory(self, addr):
        pi = addr >> 12
        page = self.pages.get(pi)
        if page is None:
            self.rtrace[pi] = bytes(PAGE)
            self.set_hash_trace(pi, 0)
            return 0
        if pi not in self.rtrace:
            self.rtrace[pi] = bytes(page)
            self.set_hash_trace(pi, 0)
        return int.from_bytes(page[addr & 0xFFF:(addr & 0xFFF) + 4], "big")

    def set_memory(self, addr, v):
        pi = addr >> 12
        page = self.pages.get(pi)
        if page is None:
            page = self.pages[pi] = bytearray(PAGE)
        if pi not in self.rtrace:
            self.rtrace[pi] = bytes(page)
            self.set_hash_trace(pi, 0)
        self.wtrace[0][pi] = page
        page[addr & 0xFFF:(addr & 0xFFF) + 4] = v.to_bytes(4, "big")

    def set_hash_range(self, page_index, digest, level):
        hash_addr = (page_index << 5) + MAX_MEMORY
        hp, off = hash_addr >> 12, hash_addr & 0xFFF
        page = self.pages.get(hp)
        if page is None:
            page = self.alloc_hash_page(hp, level)
        page[off:off + 32] = digest
        if level < 2:
            self.wtrace[level + 1][hp] = page

    def update_page_hash(self):
        for level in range(3):
            for pi in sorted(self.wtrace[level]):
                self.set_hash_range(pi, hash_page(self.wtrace[level][pi]), level)
            self.wtrace[level].clear()

    def compute_image_id(self, pc, registers):
        page = self.pages.get(ROOT_PAGE)
        if page is None:
            raise RuntimeError("compute image ID fail")
        page[REGISTERS_OFFSET:REGISTERS_OFFSET + len(registers)] = registers
        root = hash_page(page)
        return hashlib.blake2s(root + pc.to_bytes(4, "little")).digest(), root

    def get_input_image(self):
        image = {pi: data for pi, data in sorted(self.rtrace.items())}
        self.rtrace.clear()
        return image


def split_reference(mem, pc, registers):
    mem.update_page_hash()
    mem.get_memory(INIT_SP)
    mem.get_memory(INIT_SP + PAGE)
    image_id, root = mem.compute_image_id(pc, registers)
    return mem.get_input_image(), image_id, root


class TreeModel:
    """zkm_pagetree_t as include/zkm_b200.h documents it: the hash pages at and above MAX_MEMORY, absent pages read as the constant
    page of their level; split = update_page_hash over the given dirty pages followed by compute_image_id."""
    def __init__(self):
        self.pages = {}

    def _set(self, page_index, digest, level, dirty):
        hash_addr = (page_index << 5) + MAX_MEMORY
        hp, off = hash_addr >> 12, hash_addr & 0xFFF
        page = self.pages.setdefault(hp, bytearray(CONST[level]))
        page[off:off + 32] = digest
        if level < 2:
            dirty[level + 1].add(hp)

    def split(self, dirty_pages, pc, registers):
        dirty = [None, set(), set()]
        for pi, data in dirty_pages:
            assert pi < 0x80000
            self._set(pi, hash_page(data), 0, dirty)
        for level in (1, 2):
            for hp in sorted(dirty[level]):
                self._set(hp, hash_page(self.pages[hp]), level, dirty)
        page = self.pages.get(ROOT_PAGE)
        if page is None:
            raise RuntimeError("compute image ID fail")
        page[REGISTERS_OFFSET:REGISTERS_OFFSET + len(registers)] = registers
        root = hash_page(page)
        return hashlib.blake2s(root + pc.to_bytes(4, "little")).digest(), root

    def page(self, hp):
        return self.pages.get(hp)


def split_adaptor(mem, tree, pc, registers):
    """shim/src/b200_split.rs, split_segment_b200."""
    mem.get_memory(INIT_SP)
    mem.get_memory(INIT_SP + PAGE)
    dirty = [(pi, bytes(mem.wtrace[0][pi])) for pi in sorted(mem.wtrace[0])]
    image = {pi: data for pi, data in sorted(mem.rtrace.items())}
    image_id, root = tree.split(dirty, pc, registers)
    touched = sorted({0x80000 + (pi >> 7) for pi, _ in dirty} | {0x81000 + (pi >> 14) for pi, _ in dirty} | {ROOT_PAGE})
    for hp in touched:
        data = tree.page(hp)
        if data is not None:
            if hp in mem.pages:
                mem.pages[hp][:] = data
            else:
                mem.pages[hp] = bytearray(data)
    mem.wtrace[0].clear()
    mem.rtrace.clear()
    return image, image_id, root


def _program(rng, n_ops, code_page):
    """Random loads and stores; every step fetches an instruction first, as mips_step does."""
    regions = [0x10000, 0x10000 + (1 << 19), 0x400000, 0x40000000, INIT_SP - 3 * PAGE, 0x7FFF0000]
    for _ in range(n_ops):
        yield "r", code_page + 4 * int(rng.integers(0, 1024))
        base = regions[int(rng.integers(0, len(regions)))]
        addr = base + PAGE * int(rng.integers(0, 6)) + 4 * int(rng.integers(0, 1024))
        yield ("w" if rng.random() < 0.4 else "r"), addr


@pytest.mark.parametrize("seed", range(6))
def test_adaptor_order_gives_the_reference_segment_files(seed):
    rng = np.random.default_rng(seed)
    ref, ada, tree = RefMemory(), RefMemory(), TreeModel()
    code_page = 0x1000
    for m in (ref, ada):                                      # the loaded program (load_elf writes through set_memory)
        for k in range(64):
            m.set_memory(code_page + 4 * k, 0x1000 + k)
    regs0 = bytes(156)
    a = split_reference(ref, 0x1000, regs0)                   # split_prog_into_segs: one proof = false call before the first step
    b = split_adaptor(ada, tree, 0x1000, regs0)
    assert a[1:] == b[1:]                                     # ids agree; this first image is discarded upstream (and may differ)
    for seg in range(5):
        ops = list(_program(rng, int(rng.integers(1, 60)), code_page))
        if seg == 3:
            ops = ops[:1]                                     # a segment that only fetches: no dirty page at all
        for kind, addr in ops:
            v = int(rng.integers(0, 1 << 32))
            for m in (ref, ada):
                m.set_memory(addr, v) if kind == "w" else m.get_memory(addr)
        regs = bytes(rng.integers(0, 256, size=156, dtype=np.uint8))
        pc = 0x1000 + 4 * seg
        a = split_reference(ref, pc, regs)
        b = split_adaptor(ada, tree, pc, regs)
        assert a[1:] == b[1:], seg
        assert a[0].keys() == b[0].keys(), seg
        for pi in a[0]:
            assert a[0][pi] == b[0][pi], (seg, hex(pi))
        assert ROOT_PAGE in a[0] and (INIT_SP >> 12) in a[0]
        # the emulator's memory afterwards: same pages, same contents (the next segment records hash pages from it)
        assert ref.pages.keys() == ada.pages.keys()
        for pi in ref.pages:
            assert bytes(ref.pages[pi]) == bytes(ada.pages[pi]), (seg, hex(pi))
        assert not ada.rtrace and not any(ada.wtrace) and not ref.rtrace and not any(ref.wtrace)


def test_the_order_matters_without_the_argument():
    """The same adaptor with the INIT_SP reads AFTER the device call would record the root page with the new registers: the model
    does distinguish the orders, i.e. the equality above is not vacuous."""
    ref, ada, tree = RefMemory(), RefMemory(), TreeModel()
    for m in (ref, ada):
        m.set_memory(0x1000, 7)
    split_reference(ref, 0, bytes(156))
    split_adaptor(ada, tree, 0, bytes(156))
    # a segment that touches nothing before the boundary (not reachable upstream: every step fetches) shows the difference
    regs = bytes(range(156))
    a = split_reference(ref, 4, regs)
    image_id, root = tree.split([], 4, regs)
    ada.pages[ROOT_PAGE][:] = tree.page(ROOT_PAGE)
    ada.get_memory(INIT_SP)
    assert (image_id, root) == a[1:]
    assert ada.rtrace[ROOT_PAGE] != a[0][ROOT_PAGE]
