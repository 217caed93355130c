//! One device + one stream (feo_ctx).  No CPU fallback: creation fails without a CUDA device.
use crate::ffi;
use std::ffi::CStr;
use std::rc::Rc;

pub struct Context {
    pub(crate) raw: *mut ffi::feo_ctx,
}
// The C ABI is "one host thread per context": feo_ctx keeps its last error, launch window and workspace without a lock.
// So a Context is neither Send nor Sync (the raw pointer already makes it so, and `Rc` keeps every tensor that shares it on
// the thread that created it): safe Rust cannot reach the same feo_ctx from two threads.  Each thread gets its own default.

impl Context {
    pub fn new(device: i32) -> Result<Rc<Context>, String> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { ffi::feo_init(device, &mut raw) };
        if rc != ffi::FEO_OK {
            return Err(format!("feo_init(device={device}) failed with status {rc}: no usable CUDA device"));
        }
        Ok(Rc::new(Context { raw }))
    }

    /// This thread's default context on device 0 (created on first use).
    pub fn current() -> Rc<Context> {
        thread_local! {
            static CURRENT: Rc<Context> = Context::new(0).expect("feotensor-b200 needs a CUDA device");
        }
        CURRENT.with(|c| c.clone())
    }

    /// Integer x/0 and MIN/-1 panic in the reference; the kernels write 0 and raise a sticky flag that `feo_sync` /
    /// `feo_download` turn into FEO_ERR_ARITH (-> panic in `check`).  This polls it explicitly.
    pub fn arith_flag(&self, clear: bool) -> bool {
        let mut f: std::os::raw::c_int = 0;
        self.check(unsafe { ffi::feo_arith_flag(self.raw, &mut f, clear as std::os::raw::c_int) });
        f != 0
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
