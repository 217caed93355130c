//! One device + one stream (feo_ctx).  No CPU fallback: creation fails without a CUDA device.
use crate::ffi;
use std::ffi::CStr;
use std::sync::{Arc, OnceLock};

pub struct Context {
    pub(crate) raw: *mut ffi::feo_ctx,
}
// The C ABI asks for one host thread per context at a time; `Arc<Context>` is shared by the tensors it owns.
unsafe impl Send for Context {}
unsafe impl Sync for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Arc<Context>, String> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { ffi::feo_init(device, &mut raw) };
        if rc != ffi::FEO_OK {
            return Err(format!("feo_init(device={device}) failed with status {rc}: no usable CUDA device"));
        }
        Ok(Arc::new(Context { raw }))
    }

    /// Process-wide default context on device 0 (created on first use).
    pub fn current() -> Arc<Context> {
        static CURRENT: OnceLock<Arc<Context>> = OnceLock::new();
        CURRENT.get_or_init(|| Context::new(0).expect("feotensor-b200 needs a CUDA device")).clone()
    }

    pub fn sync(&self) {
        self.check(unsafe { ffi::feo_sync(self.raw) });
    }

    /// Operators have no error channel in the reference API: a backend failure panics with the library's message.
    pub(crate) fn check(&self, rc: i32) {
        if rc != ffi::FEO_OK {
            let msg = unsafe { CStr::from_ptr(ffi::feo_last_error(self.raw)) }.to_string_lossy().into_owned();
            panic!("feotensor_b200: status {rc}: {msg}");
        }
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::feo_shutdown(self.raw) };
    }
}
