//! `emulator/src/b200_split.rs` -- `InstrumentedState::split_segment` with the page hashing on the device.
//!
//! SOURCE ONLY (no cargo in this image), written against zkMIPS/zkm @ 04117ce3.  Replaces the body of `split_segment`
//! (emulator/src/state.rs:1477-1530): `Memory::update_page_hash` + `compute_image_id` (memory.rs:415-471, 129 Poseidon
//! permutations per dirty 4 KiB page, on one core) and `serde_json::to_vec(&segment)` become one call of
//! `zkm_b200_splitter_split` (include/zkm_b200.h).  `shim/emulator_b200.patch` makes the rest of the crate ready for it: `pub(crate)`
//! on `Memory::{pages, rtrace, wtrace}` (memory.rs:121-136) and on the `pre_*` fields of `InstrumentedState` (state.rs:552-560), a
//! `splitter: *mut c_void` field created in `InstrumentedState::new` with `splitter_create`, and `split_segment` forwarding here
//! under the `b200` feature.  A state loaded from a segment file calls `splitter_seed` before its first split; the host binary
//! calls `zkm_b200_init(0, ..)` once per process.
//!
//! What stays on the host and why the order below reproduces the reference's files byte for byte:
//!  * `rtrace` (the segment's memory image) also holds hash pages: every first touch of a page records the L1 / L2 / root hash
//!    pages above it as they are AT THAT MOMENT (`set_hash_trace`, memory.rs:203-219), i.e. before this boundary's update.  The
//!    emulator's own copies of the hash pages must therefore stay current: after the device call the pages the update touched are
//!    copied back into `Memory::pages` (`zkm_b200_pagetree_page`).
//!  * The reference reads `INIT_SP` and `INIT_SP + PAGE_SIZE` between `update_page_hash` and `compute_image_id` (state.rs:1489-1491).
//!    Here they are read BEFORE the device call, which does both steps at once.  The recorded pages are the same: a hash page that
//!    this boundary's update changes has a dirty page below it, and writing that page already recorded the hash page earlier in the
//!    segment; a hash page first recorded by these two reads is one the update leaves as it was.  (The root page also takes the
//!    registers in `compute_image_id`, but any segment that executed an instruction has recorded it at its first fetch; only the
//!    initial `proof = false` call differs, and its image is thrown away.)
//!  * `wtrace[1]`, `wtrace[2]` are bookkeeping of the host-side tree walk and stay empty.
//! `tests/test_page_hash.py::test_splitter_follows_split_segment` drives the same entry point against a model of the reference.
use std::ffi::{c_char, c_int, c_void, CStr};
use std::io::Write;

use crate::memory::INIT_SP;
use crate::page::{CachedPage, PAGE_SIZE};
use crate::state::{InstrumentedState, PAGE_CYCLES};

#[repr(C)]
pub struct ZkmSplitState {
    pub dirty_page_indices: *const u32,
    pub dirty_pages: *const u8,
    pub n_dirty_pages: usize,
    pub read_page_indices: *const u32,
    pub read_pages: *const u8,
    pub n_read_pages: usize,
    pub registers: *const u8,
    pub pc: u32,
    pub step: u64,
    pub input_stream: *const *const u8,
    pub input_stream_lens: *const usize,
    pub n_input_streams: usize,
    pub input_stream_ptr: u64,
    pub public_values_stream: *const u8,
    pub public_values_stream_len: usize,
    pub public_values_stream_ptr: u64,
}

extern "C" {
    fn zkm_b200_splitter_create(out: *mut *mut c_void, err: *mut *mut c_char) -> c_int;
    fn zkm_b200_splitter_destroy(s: *mut c_void);
    fn zkm_b200_splitter_pagetree(s: *mut c_void) -> *mut c_void;
    fn zkm_b200_splitter_split(
        s: *mut c_void, state: *const ZkmSplitState, proof: c_int, segment_json_out: *mut *mut c_char, segment_json_len: *mut usize,
        image_id_out: *mut u8, page_hash_root_out: *mut u8, err: *mut *mut c_char,
    ) -> c_int;
    fn zkm_b200_pagetree_page(t: *const c_void, page_index: u32, out: *mut u8, present: *mut c_int, err: *mut *mut c_char) -> c_int;
    fn zkm_b200_pagetree_set_page(t: *mut c_void, page_index: u32, data: *const u8, err: *mut *mut c_char) -> c_int;
    fn zkm_b200_free_string(s: *mut c_char);
}

fn check(rc: c_int, err: *mut c_char) {
    if rc != 0 {
        let msg = if err.is_null() { "zkm_b200: unknown error".to_string() } else { unsafe { CStr::from_ptr(err) }.to_string_lossy().into_owned() };
        unsafe { zkm_b200_free_string(err) };
        panic!("{msg}"); // split_segment has no error channel upstream either ("compute image ID fail" is a panic there)
    }
}

pub fn splitter_create() -> *mut c_void {
    let (mut s, mut err) = (core::ptr::null_mut(), core::ptr::null_mut());
    check(unsafe { zkm_b200_splitter_create(&mut s, &mut err) }, err);
    s
}

/// A state resumed from a segment file (`State::load_seg`, as `split_seg_into_segs` does, utils.rs:62-109) has its hash pages in
/// memory already: hand them to the library's tree before the first split.  Not needed for `split_prog_into_segs`.
pub fn splitter_seed(s: *mut c_void, memory: &crate::memory::Memory) {
    let tree = unsafe { zkm_b200_splitter_pagetree(s) };
    for (index, page) in memory.pages.range(0x80000u32..=0x81020u32) {
        let mut err = core::ptr::null_mut();
        check(unsafe { zkm_b200_pagetree_set_page(tree, *index, page.borrow().data.as_ptr(), &mut err) }, err);
    }
}

/// `InstrumentedState` cannot implement `Drop` (split_prog_into_segs moves `state` out of it, utils.rs:51-55): call this where the
/// state is given up, or leave the one splitter of a run to process exit.
pub fn splitter_destroy(s: *mut c_void) {
    unsafe { zkm_b200_splitter_destroy(s) }
}

/// BTreeMap order = ascending page index, which is what the library requires.
fn gather(pages: impl Iterator<Item = (u32, [u8; PAGE_SIZE])>) -> (Vec<u32>, Vec<u8>) {
    let (mut idx, mut bytes) = (Vec::new(), Vec::new());
    for (i, data) in pages {
        idx.push(i);
        bytes.extend_from_slice(&data);
    }
    (idx, bytes)
}

impl InstrumentedState {
    pub fn split_segment_b200<W: Write>(&mut self, proof: bool, output: &str, new_writer: fn(&str) -> Option<W>) {
        self.state.total_cycle += self.state.cycle + (self.state.memory.page_count() + 1) * PAGE_CYCLES;
        self.state.total_step += self.state.step;
        let registers = self.state.get_registers_bytes();
        // load public input, assume the max size of public input is 6KB (see the module comment for the order)
        let _ = self.state.memory.get_memory(INIT_SP);
        let _ = self.state.memory.get_memory(INIT_SP + PAGE_SIZE as u32);

        let (dirty_idx, dirty) = gather(self.state.memory.wtrace[0].iter().map(|(i, p)| (*i, p.borrow().data)));
        let (read_idx, read) = gather(self.state.memory.rtrace.iter().map(|(i, d)| (*i, *d)));
        let streams: Vec<*const u8> = self.state.input_stream.iter().map(|v| v.as_ptr()).collect();
        let lens: Vec<usize> = self.state.input_stream.iter().map(|v| v.len()).collect();
        let st = ZkmSplitState {
            dirty_page_indices: dirty_idx.as_ptr(), dirty_pages: dirty.as_ptr(), n_dirty_pages: dirty_idx.len(),
            read_page_indices: read_idx.as_ptr(), read_pages: read.as_ptr(), n_read_pages: read_idx.len(),
            registers: registers.as_ptr(), pc: self.state.pc, step: self.state.step,
            input_stream: streams.as_ptr(), input_stream_lens: lens.as_ptr(), n_input_streams: streams.len(),
            input_stream_ptr: self.state.input_stream_ptr as u64,
            public_values_stream: self.state.public_values_stream.as_ptr(), public_values_stream_len: self.state.public_values_stream.len(),
            public_values_stream_ptr: self.state.public_values_stream_ptr as u64,
        };
        let (mut text, mut len, mut err): (*mut c_char, usize, *mut c_char) = (core::ptr::null_mut(), 0, core::ptr::null_mut());
        let (mut image_id, mut page_hash_root) = ([0u8; 32], [0u8; 32]);
        check(
            unsafe { zkm_b200_splitter_split(self.splitter, &st, proof as c_int, &mut text, &mut len, image_id.as_mut_ptr(), page_hash_root.as_mut_ptr(), &mut err) },
            err,
        );

        // the hash pages this boundary changed, back into the emulator's memory: L1 = 0x80000 + (i >> 7), L2 = 0x81000 + (i >> 14),
        // root = 0x81020 (set_hash_range, memory.rs:388-413: hash_addr = (page_index << 5) + MAX_MEMORY)
        let tree = unsafe { zkm_b200_splitter_pagetree(self.splitter) };
        let mem = &mut self.state.memory;
        let mut touched: Vec<u32> = dirty_idx.iter().flat_map(|i| [0x80000 + (i >> 7), 0x81000 + (i >> 14)]).collect();
        touched.push(0x81020);
        touched.sort_unstable();
        touched.dedup();
        for hp in touched {
            let (mut data, mut present, mut err) = ([0u8; PAGE_SIZE], 0 as c_int, core::ptr::null_mut());
            check(unsafe { zkm_b200_pagetree_page(tree, hp, data.as_mut_ptr(), &mut present, &mut err) }, err);
            if present != 0 {
                let page = match mem.pages.get(&hp) {
                    Some(p) => p.clone(),
                    None => {
                        let p = std::rc::Rc::new(std::cell::RefCell::new(CachedPage::new()));
                        mem.pages.insert(hp, p.clone());
                        p
                    }
                };
                page.borrow_mut().data.copy_from_slice(&data);
            }
        }
        mem.wtrace[0].clear();
        mem.rtrace.clear(); // get_input_image() clears it (memory.rs:524-538)

        if proof {
            // the library has filled in segment_id = pre_segment_id, pc = pre_pc, pre_image_id, pre_hash_root, the pre_* streams
            let name = format!("{output}/{}", self.pre_segment_id);
            log::debug!("split: file {}", name);
            let mut f = new_writer(&name).unwrap();
            f.write_all(unsafe { std::slice::from_raw_parts(text as *const u8, len) }).unwrap();
            unsafe { zkm_b200_free_string(text) };
            self.pre_segment_id += 1;
        }
        // mirrored on the host because other code reads them (`pre_segment_id` is returned by split_prog_into_segs, utils.rs:51-55)
        self.pre_input = self.state.input_stream.clone();
        self.pre_input_ptr = self.state.input_stream_ptr;
        self.pre_public_values = self.state.public_values_stream.clone();
        self.pre_public_values_ptr = self.state.public_values_stream_ptr;
        self.pre_pc = self.state.pc;
        self.pre_image_id = image_id;
        self.pre_hash_root = page_hash_root;
        self.state.cycle = 0;
        self.state.step = 0;
    }
}
