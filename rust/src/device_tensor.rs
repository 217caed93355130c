//! `DeviceTensor<T>`: the reference's `DynamicTensor<T>` (feotensor src/tensor.rs:15-505) with device-resident storage.
//! Same constructors, accessors, reductions, `prod`, `pow` and `std::ops` impls; the arithmetic runs in
//! libfeotensor_b200.so.  Where the reference returns `Err(ShapeError)` so does this; where it panics so does this.
use crate::context::Context;
use crate::device_matrix::DeviceMatrix;
use crate::device_storage::{dims_of, DeviceElem, DeviceStorage};
use crate::device_vector::DeviceVector;
use crate::ffi;
use feotensor::axes::Axes;
use feotensor::coordinate::Coordinate;
use feotensor::error::ShapeError;
use feotensor::shape::Shape;
use num::Float;
use std::ops::{Add, Div, Mul, Sub};
use std::os::raw::c_int;

pub struct DeviceTensor<T: DeviceElem> {
    pub(crate) data: DeviceStorage<T>,
    pub(crate) shape: Shape,
}

impl<T: DeviceElem> DeviceTensor<T> {
    /// tensor.rs:22-30
    pub fn new(shape: &Shape, data: &[T]) -> Result<DeviceTensor<T>, ShapeError> {
        if data.len() != shape.size() {
            return Err(ShapeError::new("Data length does not match shape size"));
        }
        Ok(DeviceTensor { data: DeviceStorage::new(&Context::current(), data.to_vec()), shape: shape.clone() })
    }

    pub(crate) fn uninit(ctx: &std::rc::Rc<Context>, shape: &Shape) -> DeviceTensor<T> {
        DeviceTensor { data: DeviceStorage::uninit(ctx, shape.size()), shape: shape.clone() }
    }

    /// tensor.rs:32-38
    pub fn fill(shape: &Shape, value: T) -> DeviceTensor<T> {
        let t = Self::uninit(&Context::current(), shape);
        t.data.ctx.check(unsafe { ffi::feo_fill(t.data.ctx.raw, T::DTYPE, t.data.ptr, (&value as *const T).cast(), shape.size()) });
        t
    }
    pub fn zeros(shape: &Shape) -> DeviceTensor<T> { Self::fill(shape, T::zero()) }
    pub fn ones(shape: &Shape) -> DeviceTensor<T> { Self::fill(shape, T::one()) }

    pub fn shape(&self) -> &Shape { &self.shape }
    pub fn size(&self) -> usize { self.shape.size() }

    /// tensor.rs:54-56 (by value: the element lives on the device)
    pub fn get(&self, coord: &Coordinate) -> Result<T, ShapeError> {
        Ok(self.data.get(self.data.flatten(coord, &self.shape)?))
    }
    /// tensor.rs:63-67
    pub fn set(&mut self, coord: &Coordinate, value: T) -> Result<(), ShapeError> {
        let index = self.data.flatten(coord, &self.shape)?;
        self.data.set(index, value);
        Ok(())
    }
    pub fn to_vec(&self) -> Vec<T> { self.data.to_vec() }

    fn reduce(&self, kind: c_int, axes: Axes) -> DeviceTensor<T> {
        let dims = dims_of(&self.shape);
        let (mut out_rank, mut out_dims) = (0 as c_int, vec![0usize; dims.len().max(1)]);
        let rc = unsafe {
            ffi::feo_reduce_shape(dims.len() as c_int, dims.as_ptr(), axes.len() as c_int, axes.as_ptr(), &mut out_rank, out_dims.as_mut_ptr())
        };
        // the reference panics on axes that are not strictly ascending / in range (tensor.rs:97-100 with coordinate.rs:27-33)
        assert!(rc == ffi::FEO_OK, "axes must be strictly ascending, unique and < order");
        out_dims.truncate(out_rank as usize);
        let out = Self::uninit(&self.data.ctx, &Shape::new(out_dims).unwrap());
        self.data.ctx.check(unsafe {
            ffi::feo_reduce(self.data.ctx.raw, kind, T::DTYPE, out.data.ptr, self.data.ptr, dims.len() as c_int, dims.as_ptr(),
                            axes.len() as c_int, axes.as_ptr())
        });
        out
    }
    pub fn sum(&self, axes: Axes) -> DeviceTensor<T> { self.reduce(ffi::FEO_SUM, axes) }   // tensor.rs:70-108
    pub fn mean(&self, axes: Axes) -> DeviceTensor<T> { self.reduce(ffi::FEO_MEAN, axes) } // tensor.rs:110-132
    pub fn var(&self, axes: Axes) -> DeviceTensor<T> { self.reduce(ffi::FEO_VAR, axes) }   // tensor.rs:134-203
    pub fn max(&self, axes: Axes) -> DeviceTensor<T> { self.reduce(ffi::FEO_MAX, axes) }   // tensor.rs:205-255
    pub fn min(&self, axes: Axes) -> DeviceTensor<T> { self.reduce(ffi::FEO_MIN, axes) }   // tensor.rs:257-307

    /// Tensor product, consistent with numpy.tensordot(a, b, axis=0) (tensor.rs:309-322)
    pub fn prod(&self, other: &DeviceTensor<T>) -> DeviceTensor<T> {
        let out = Self::uninit(&self.data.ctx, &self.shape.stack(&other.shape));
        self.data.ctx.check(unsafe {
            ffi::feo_outer(self.data.ctx.raw, T::DTYPE, out.data.ptr, self.data.ptr, self.size(), other.data.ptr, other.size())
        });
        out
    }

    /// Broadcasting elementwise op: NEW surface (the reference asserts equal shapes, tensor.rs:366).
    pub fn broadcast(&self, op: c_int, rhs: &DeviceTensor<T>) -> Result<DeviceTensor<T>, ShapeError> {
        let (a, b) = (dims_of(&self.shape), dims_of(&rhs.shape));
        if a.len() != b.len() {
            return Err(ShapeError::new(&format!("incorrect order ({} vs {}).", a.len(), b.len())));
        }
        let mut od = vec![0usize; a.len()];
        for d in 0..a.len() {
            if a[d] != b[d] && a[d] != 1 && b[d] != 1 {
                return Err(ShapeError::new("shapes are not broadcastable"));
            }
            od[d] = a[d].max(b[d]);
        }
        let (mut sa, mut sb, mut xa, mut xb) = (vec![0usize; a.len()], vec![0usize; a.len()], 1usize, 1usize);
        for d in (0..a.len()).rev() {
            sa[d] = if a[d] == od[d] && a[d] != 1 { xa } else { 0 };
            sb[d] = if b[d] == od[d] && b[d] != 1 { xb } else { 0 };
            xa *= a[d];
            xb *= b[d];
        }
        let out = Self::uninit(&self.data.ctx, &Shape::new(od.clone())?);
        self.data.ctx.check(unsafe {
            ffi::feo_ewise_broadcast(self.data.ctx.raw, op, T::DTYPE, out.data.ptr, self.data.ptr, rhs.data.ptr, od.len() as c_int,
                                     od.as_ptr(), sa.as_ptr(), sb.as_ptr())
        });
        Ok(out)
    }

    pub(crate) fn binary(self, op: c_int, rhs: DeviceTensor<T>) -> DeviceTensor<T> {
        assert!(self.shape == rhs.shape); // tensor.rs:366, 409, 439, 482
        let out = Self::uninit(&self.data.ctx, &self.shape);
        self.data.ctx.check(unsafe { ffi::feo_ewise(self.data.ctx.raw, op, T::DTYPE, out.data.ptr, self.data.ptr, rhs.data.ptr, self.size()) });
        out // `self` and `rhs` drop here: feo_free is stream-ordered, i.e. after the kernel that reads them
    }
    pub(crate) fn scalar(self, op: c_int, rhs: T) -> DeviceTensor<T> {
        let out = Self::uninit(&self.data.ctx, &self.shape);
        self.data.ctx.check(unsafe {
            ffi::feo_ewise_scalar(self.data.ctx.raw, op, T::DTYPE, out.data.ptr, self.data.ptr, (&rhs as *const T).cast(), self.size())
        });
        out
    }
}

impl<T: DeviceElem + Float> DeviceTensor<T> {
    /// tensor.rs:325-333
    pub fn pow(&self, power: T) -> DeviceTensor<T> {
        let out = Self::uninit(&self.data.ctx, &self.shape);
        self.data.ctx.check(unsafe { ffi::feo_pow(self.data.ctx.raw, T::DTYPE, out.data.ptr, self.data.ptr, (&power as *const T).cast(), self.size()) });
        out
    }
}

// The sixteen operator impls of tensor.rs:336-505.  Tensor (op) Vector / Matrix return the newtype like the reference;
// `Tensor - Vector` is computed directly (the reference's `(rhs * (0-1)) + self`, tensor.rs:418-424, is bit-identical for
// floats and wrapping integers); `Tensor / Vector` unwraps `from_tensor(self)` first and so panics unless order == 1.
macro_rules! tensor_ops {
    ($tr:ident, $f:ident, $code:expr) => {
        impl<T: DeviceElem> $tr<T> for DeviceTensor<T> {
            type Output = DeviceTensor<T>;
            fn $f(self, rhs: T) -> DeviceTensor<T> { self.scalar($code, rhs) }
        }
        impl<T: DeviceElem> $tr<DeviceTensor<T>> for DeviceTensor<T> {
            type Output = DeviceTensor<T>;
            fn $f(self, rhs: DeviceTensor<T>) -> DeviceTensor<T> { self.binary($code, rhs) }
        }
        impl<T: DeviceElem> $tr<DeviceVector<T>> for DeviceTensor<T> {
            type Output = DeviceVector<T>;
            fn $f(self, rhs: DeviceVector<T>) -> DeviceVector<T> { DeviceVector::from_tensor(self.binary($code, rhs.into_tensor())).unwrap() }
        }
        impl<T: DeviceElem> $tr<DeviceMatrix<T>> for DeviceTensor<T> {
            type Output = DeviceMatrix<T>;
            fn $f(self, rhs: DeviceMatrix<T>) -> DeviceMatrix<T> { DeviceMatrix::from_tensor(self.binary($code, rhs.into_tensor())).unwrap() }
        }
    };
}
tensor_ops!(Add, add, ffi::FEO_ADD);
tensor_ops!(Sub, sub, ffi::FEO_SUB);
tensor_ops!(Mul, mul, ffi::FEO_MUL);
tensor_ops!(Div, div, ffi::FEO_DIV);
