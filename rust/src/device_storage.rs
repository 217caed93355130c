//! `DeviceStorage<T>`: the device-resident type that sits beside `DynamicStorage<T>` (feotensor src/storage.rs:7-100).
use crate::context::Context;
use crate::ffi;
use feotensor::coordinate::Coordinate;
use feotensor::error::ShapeError;
use feotensor::shape::Shape;
use std::marker::PhantomData;
use std::os::raw::{c_int, c_void};
use std::rc::Rc;

/// Element types the backend knows (tensor.rs:21 admits any `Num + PartialOrd + Copy` primitive).
pub unsafe trait DeviceElem: num::Num + PartialOrd + Copy {
    const DTYPE: c_int;
}
unsafe impl DeviceElem for f32 { const DTYPE: c_int = ffi::FEO_F32; }
unsafe impl DeviceElem for f64 { const DTYPE: c_int = ffi::FEO_F64; }
unsafe impl DeviceElem for i32 { const DTYPE: c_int = ffi::FEO_I32; }
unsafe impl DeviceElem for i64 { const DTYPE: c_int = ffi::FEO_I64; }
unsafe impl DeviceElem for u32 { const DTYPE: c_int = ffi::FEO_U32; }

pub(crate) fn dims_of(shape: &Shape) -> Vec<usize> {
    (0..shape.order()).map(|i| shape[i]).collect()
}

/// Dense row-major device buffer; indexing is `DynamicStorage::flatten` (storage.rs:18-38).
pub struct DeviceStorage<T: DeviceElem> {
    pub(crate) ctx: Rc<Context>,
    pub(crate) ptr: *mut c_void,
    len: usize,
    _t: PhantomData<T>,
}

impl<T: DeviceElem> DeviceStorage<T> {
    pub(crate) fn uninit(ctx: &Rc<Context>, len: usize) -> Self {
        let mut ptr = std::ptr::null_mut();
        ctx.check(unsafe { ffi::feo_malloc(ctx.raw, len * std::mem::size_of::<T>(), &mut ptr) });
        DeviceStorage { ctx: ctx.clone(), ptr, len, _t: PhantomData }
    }

    /// `DynamicStorage::new(data)` (storage.rs:13-15): copies the host data to the device.
    pub fn new(ctx: &Rc<Context>, data: Vec<T>) -> Self {
        let s = Self::uninit(ctx, data.len());
        ctx.check(unsafe { ffi::feo_upload(ctx.raw, s.ptr, data.as_ptr().cast(), data.len() * std::mem::size_of::<T>()) });
        ctx.sync(); // `data` is dropped on return: the pageable copy must have left it
        s
    }

    /// storage.rs:18-38, same two error cases and messages.
    pub fn flatten(&self, coord: &Coordinate, shape: &Shape) -> Result<usize, ShapeError> {
        let c: Vec<usize> = coord.iter().copied().collect();
        let d = dims_of(shape);
        let (mut idx, mut detail) = (0usize, 0 as c_int);
        let rc = unsafe { ffi::feo_flatten(c.len() as c_int, c.as_ptr(), d.len() as c_int, d.as_ptr(), &mut idx, &mut detail) };
        if rc == ffi::FEO_OK {
            Ok(idx)
        } else if detail < 0 {
            Err(ShapeError::new(&format!("incorrect order ({} vs {}).", coord.order(), shape.order())))
        } else {
            Err(ShapeError::new(&format!("out of bounds for dimension {}", detail)))
        }
    }

    pub fn size(&self) -> usize {
        self.len
    }

    pub fn get(&self, index: usize) -> T {
        assert!(index < self.len, "index out of bounds");
        let mut v = T::zero();
        let src = unsafe { (self.ptr as *const u8).add(index * std::mem::size_of::<T>()) };
        self.ctx.check(unsafe { ffi::feo_download(self.ctx.raw, (&mut v as *mut T).cast(), src.cast(), std::mem::size_of::<T>()) });
        v
    }

    pub fn set(&mut self, index: usize, value: T) {
        assert!(index < self.len, "index out of bounds");
        let dst = unsafe { (self.ptr as *mut u8).add(index * std::mem::size_of::<T>()) };
        self.ctx.check(unsafe { ffi::feo_upload(self.ctx.raw, dst.cast(), (&value as *const T).cast(), std::mem::size_of::<T>()) });
        self.ctx.sync();
    }

    /// Download everything (the reference hands out iterators over its Vec, storage.rs:40-100).
    pub fn to_vec(&self) -> Vec<T> {
        let mut v = vec![T::zero(); self.len];
        self.ctx.check(unsafe { ffi::feo_download(self.ctx.raw, v.as_mut_ptr().cast(), self.ptr, self.len * std::mem::size_of::<T>()) });
        v
    }
}

impl<T: DeviceElem> Drop for DeviceStorage<T> {
    /// Stream-ordered free: safe immediately after the op that consumed this buffer has been enqueued.
    fn drop(&mut self) {
        unsafe { ffi::feo_free(self.ctx.raw, self.ptr) };
    }
}
