// Bodies for the patched plonky2 (fork of rev 666f3151) -- what replaces the CPU code of
// plonky2/src/fri/oracle.rs and plonky2/src/hash/merkle_tree.rs.  NOT compiled in this repository (no Rust toolchain
// here); written against plonky2-b200-sys, whose declarations are generated from include/plonky2_b200.h.
//
// The reference itself does not change: eth-lc-plonky2/src/main.rs:227 (`builder.build::<C>()`) and :230
// (`data.prove(witness)`) reach these through the [patch] entry of the workspace Cargo.toml.

use plonky2_b200_sys as sys;
use std::ptr;

/// Owner of a device-resident batch; `Drop` releases it (the C side defers the free in stream order).
pub struct DeviceBatch(pub *mut sys::eng_batch);
unsafe impl Send for DeviceBatch {}
unsafe impl Sync for DeviceBatch {}
impl Drop for DeviceBatch {
    fn drop(&mut self) {
        unsafe { sys::eng_batch_free(self.0) };
    }
}

impl<F: RichField + Extendable<D>, C: GenericConfig<D, F = F>, const D: usize> PolynomialBatch<F, C, D> {
    /// plonky2::fri::oracle::PolynomialBatch::from_values -- "IFFT", "FFT + blinding", "transpose LDEs" and "build Merkle
    /// tree" happen in one call on the device; `timing` and `fft_root_table` are ignored.
    pub fn from_values(values: Vec<PolynomialValues<F>>, rate_bits: usize, blinding: bool, cap_height: usize,
                       _timing: &mut TimingTree, _fft_root_table: Option<&FftRootTable<F>>) -> Self {
        let n = values[0].len();
        assert!(values.iter().all(|v| v.len() == n), "All polynomials must have the same length");
        let degree_log = log2_strict(n);
        // GoldilocksField is #[repr(transparent)] over u64: the columns go across as they are (pageable memory is staged
        // through pinned bounce buffers inside the engine and overlapped with the transforms)
        let cols: Vec<*const u64> = values.iter().map(|v| v.values.as_ptr() as *const u64).collect();
        let mut h: *mut sys::eng_batch = ptr::null_mut();
        sys::check(unsafe {
            sys::eng_batch_from_values(cols.as_ptr(), cols.len() as u32, degree_log as u32, rate_bits as u32,
                                       blinding as i32, rand::random::<u64>(), cap_height as u32, &mut h)
        });
        Self::from_device(DeviceBatch(h), degree_log, rate_bits, blinding, cap_height)
    }

    fn from_device(dev: DeviceBatch, degree_log: usize, rate_bits: usize, blinding: bool, cap_height: usize) -> Self {
        let mut cap = vec![[0u64; 4]; 1 << cap_height];
        sys::check(unsafe { sys::eng_batch_cap(dev.0, cap.as_mut_ptr() as *mut u64) });
        Self {
            polynomials: LazyCoeffs::new(&dev),           // eng_batch_coeffs on first host access (OpeningSet::new)
            merkle_tree: MerkleTree::from_device(&dev, cap, cap_height),   // get/prove -> eng_batch_leaves / eng_batch_merkle_path
            degree_log,
            rate_bits,
            blinding,
            dev,
        }
    }

    /// get_lde_values(index, step) = leaves[bitrev(index * step)] without the salt
    pub fn get_lde_values(&self, index: usize, step: usize) -> Vec<F> {
        let mut info = sys::eng_batch_info_t::default();
        sys::check(unsafe { sys::eng_batch_info(self.dev.0, &mut info) });
        let mut row = vec![F::ZERO; info.num_polys as usize];
        sys::check(unsafe { sys::eng_batch_lde_values(self.dev.0, index as u64, step as u64, row.as_mut_ptr() as *mut u64) });
        row
    }
}
