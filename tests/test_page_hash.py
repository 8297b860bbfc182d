"""The emulator's segment-splitter hashing: Memory::update_page_hash + compute_image_id,
emulator/src/memory.rs:43-125,378-471, as split_segment drives them (emulator/src/state.rs:1460-1478).

CPU: the oracle's restatement (oracle/pagehash.h) against an independent Python walk of the same reference lines built on the
Poseidon permutation alone (which the SURVEY Appendix D known answers pin).  GPU: zkm_b200_hash_pages / zkm_b200_pagetree_*
against the oracle -- every digest, every hash page, image id and root over several segments."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding


def _perm(orc, st):
    a = np.array(st, dtype=np.uint64)
    orc.orc_poseidon_permute(binding.u64ptr(a), 0)
    return [int(x) for x in a]


def py_poseidon_bytes(orc, data: bytes):
    """memory.rs:43-79, written from the reference text (not from oracle/pagehash.h)."""
    l = len(data)
    chunks = l // 32 + 1
    buf = bytearray(data) + bytearray(chunks * 32 - l)
    if l % 32 == 31:
        buf[l] = 0b10000001
    else:
        buf[l] = 1
        buf[chunks * 32 - 1] = 0b10000000
    st = [0] * 12
    for b in range(chunks):
        st[:8] = [int.from_bytes(buf[32 * b + 4 * i:32 * b + 4 * i + 4], "little") for i in range(8)]
        st = _perm(orc, st)
    return b"".join(x.to_bytes(8, "little") for x in st[:4])


class PyMemoryHashes:
    """update_page_hash / compute_image_id over a dict of hash pages (memory.rs:378-471)."""

    def __init__(self, orc):
        self.orc, self.pages = orc, {}
        self.const, h = [], py_poseidon_bytes(orc, bytes(4096))
        for _ in range(3):
            self.const.append(h * 128)
            h = py_poseidon_bytes(orc, self.const[-1])

    def set_hash_range(self, page_index, h, level):
        addr = (page_index << 5) + 0x80000000
        hp, off = addr >> 12, addr & 0xFFF
        pg = self.pages.setdefault(hp, bytearray(self.const[level]))
        pg[off:off + 32] = h
        return hp

    def split(self, dirty, registers, pc):
        w1 = sorted({self.set_hash_range(i, py_poseidon_bytes(self.orc, d), 0) for i, d in dirty})
        w2 = sorted({self.set_hash_range(i, py_poseidon_bytes(self.orc, bytes(self.pages[i])), 1) for i in w1})
        for i in w2:
            self.set_hash_range(i, py_poseidon_bytes(self.orc, bytes(self.pages[i])), 2)
        root = self.pages[0x81020]
        root[0x400:0x400 + 156] = registers
        h = py_poseidon_bytes(self.orc, bytes(root))
        fin = b"".join(h[i:i + 4][::-1] for i in range(0, 32, 4)) + pc.to_bytes(4, "little")
        return py_poseidon_bytes(self.orc, fin), h


def _segments(seed, counts):
    """Dirty-page sets of consecutive segments: sparse indices over code, heap and stack regions, some pages dirtied again."""
    rng = np.random.default_rng(seed)
    out, seen = [], []
    for n in counts:
        fresh = [int(x) for x in rng.choice(np.concatenate([np.arange(0x10, 0x90), np.arange(0x7FF00, 0x7FFFE), np.arange(0x3000, 0x3400, 7)]),
                                            size=n, replace=False)]
        again = seen[:max(0, n // 3)]
        idx = sorted(set(fresh + again))
        seen = idx
        pages = rng.integers(0, 256, size=(len(idx), 4096), dtype=np.uint8)
        if len(idx):
            pages[0, :] = 0                    # an all-zero dirty page
        regs = bytes(int(b) for b in rng.integers(0, 256, size=156))
        out.append((idx, pages, regs, int(rng.integers(0, 1 << 32))))
    return out


def test_oracle_page_hashing_matches_the_reference_walk(orc):
    rng = np.random.default_rng(3)
    for ln in (0, 1, 31, 32, 33, 36, 63, 64, 100, 4096):
        data = bytes(int(b) for b in rng.integers(0, 256, size=ln))
        out = np.zeros(4, dtype=np.uint64)
        orc.orc_poseidon_bytes(data, ln, binding.u64ptr(out))
        assert out.tobytes() == py_poseidon_bytes(orc, data), ln
    py = PyMemoryHashes(orc)
    for lv in range(3):
        pg = np.zeros(4096, dtype=np.uint8)
        orc.orc_const_hash_page(lv, pg.ctypes.data)
        assert pg.tobytes() == py.const[lv]
    tree = binding.OrcPageTree(orc)
    with pytest.raises(RuntimeError, match="compute image ID fail"):        # no page was ever hashed: the reference panics
        tree.split([], np.zeros((0, 4096), dtype=np.uint8), bytes(156), 0)
    for idx, pages, regs, pc in _segments(11, (5, 9, 1)):
        want = py.split([(i, pages[k].tobytes()) for k, i in enumerate(idx)], regs, pc)
        assert tree.split(idx, pages, regs, pc) == want
        assert tree.count() == len(py.pages)
        for hp, content in py.pages.items():
            assert tree.page(hp).tobytes() == bytes(content), hex(hp)
    tree.close()


def test_page_tree_is_seeded_without_a_device():
    """zkm_b200_pagetree_set_page / _page are host-side bookkeeping (a resumed emulator state brings its hash pages along): they
    work before zkm_b200_init, refuse main-memory indices, and a page that was never set or hashed reads as absent."""
    from zkm_b200 import lib as zl
    lib = zl.load()
    tree = zl.PageTree(lib)
    rng = np.random.default_rng(2)
    pages = {hp: rng.integers(0, 256, size=4096, dtype=np.uint8) for hp in (0x80000, 0x80FFF, 0x81000, 0x8101F, 0x81020)}
    for hp, data in pages.items():
        tree.set_page(hp, data.tobytes())
    for hp, data in pages.items():
        assert (tree.page(hp) == data).all()
    tree.set_page(0x81020, bytes(4096))                                  # overwriting is allowed
    assert not tree.page(0x81020).any() and tree.page(0x80001) is None
    for bad in (0, 0x7FFFF, 0x81021, 0xFFFFF):
        with pytest.raises(zl.ZkmError, match="not a hash page index"):
            tree.set_page(bad, bytes(4096))
    tree.close()


@pytest.mark.gpu
def test_device_page_hashing_matches_oracle(zkm, orc):
    from zkm_b200 import lib as zl
    rng = np.random.default_rng(5)
    for n in (1, 2, 33, 700):
        pages = rng.integers(0, 256, size=(n, 4096), dtype=np.uint8)
        pages[0] = 0
        want = np.zeros((n, 32), dtype=np.uint8)
        orc.orc_hash_pages(pages.ctypes.data, n, want.ctypes.data)
        assert (zl.hash_pages(zkm, pages) == want).all()
    dev, cpu = zl.PageTree(zkm), binding.OrcPageTree(orc)
    with pytest.raises(zl.ZkmError, match="compute image ID fail"):
        dev.split([], np.zeros((0, 4096), dtype=np.uint8), bytes(156), 0)
    with pytest.raises(zl.ZkmError, match="main-memory page"):
        dev.split([0x80000], np.zeros((1, 4096), dtype=np.uint8), bytes(156), 0)
    for idx, pages, regs, pc in _segments(17, (40, 300, 3, 120)):
        assert dev.split(idx, pages, regs, pc) == cpu.split(idx, pages, regs, pc)
        for hp in {0x80000 + (i >> 7) for i in idx} | {0x81000 + ((0x80000 + (i >> 7) - 0x80000) >> 7) for i in idx} | {0x81020}:
            assert (dev.page(hp) == cpu.page(hp)).all(), hex(hp)
    assert dev.page(0x12345) is None
    dev.close(); cpu.close()


@pytest.mark.gpu
def test_splitter_follows_split_segment(zkm, orc):
    """zkm_b200_splitter_split against a Python model of InstrumentedState::split_segment (emulator/src/state.rs:1477-1530) built on
    the oracle's page tree: as split_prog_into_segs drives it (utils.rs:23-57) -- one call with proof = false, then boundaries with
    proof = true.  Every segment file must be the serde_json text of the reference's Segment, image ids and roots must chain."""
    import json
    from zkm_b200 import lib as zl
    rng = np.random.default_rng(23)
    dev, cpu = zl.Splitter(zkm), binding.OrcPageTree(orc)
    pre = dict(segment_id=0, pc=0, image_id=bytes(32), hash_root=bytes(32), input=[], input_ptr=0, pv=b"", pv_ptr=0)
    streams, pv = [b"abc", bytes(range(20))], b""
    for k, (idx, pages, regs, pc) in enumerate(_segments(29, (12, 30, 7, 50))):
        proof = k > 0
        read_idx = sorted(set(idx[::2] + [0x7FFFD, 0x7FFFE]))
        read_pages = rng.integers(0, 256, size=(len(read_idx), 4096), dtype=np.uint8)
        step, in_ptr, pv_ptr = 1000 * k + 17, k, 4 * k
        pv = pv + bytes([k] * 4)
        text, image_id, root = dev.split((idx, pages), (read_idx, read_pages), regs, pc, step, streams, in_ptr, pv, pv_ptr, proof)
        want_id, want_root = cpu.split(idx, pages, regs, pc)
        assert (image_id, root) == (want_id, want_root)
        if proof:
            image = {str((pi << 12) + 4 * i): int.from_bytes(read_pages[j, 4 * i:4 * i + 4].tobytes(), "little") for j, pi in enumerate(read_idx)
                     for i in range(1024)}
            want = {"mem_image": image, "pc": pre["pc"], "segment_id": pre["segment_id"], "pre_image_id": list(pre["image_id"]),
                    "pre_hash_root": list(pre["hash_root"]), "image_id": list(want_id), "page_hash_root": list(want_root), "end_pc": pc, "step": step,
                    "input_stream": [list(b) for b in pre["input"]], "input_stream_ptr": pre["input_ptr"], "public_values_stream": list(pre["pv"]),
                    "public_values_stream_ptr": pre["pv_ptr"]}
            assert text.decode() == json.dumps(want, separators=(",", ":"))
            pre["segment_id"] += 1
        else:
            assert text is None
        pre.update(pc=pc, image_id=want_id, hash_root=want_root, input=list(streams), input_ptr=in_ptr, pv=pv, pv_ptr=pv_ptr)
        streams = streams + [bytes([k])]
    assert dev.segment_count() == 3
    dev.close(); cpu.close()
