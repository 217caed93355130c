//! `DeviceMatrix<T>`: the reference's `DynamicMatrix<T>` newtype (feotensor src/matrix.rs:13-242).
use crate::device_storage::DeviceElem;
use crate::device_tensor::DeviceTensor;
use crate::device_vector::DeviceVector;
use crate::ffi;
use feotensor::axes::Axes;
use feotensor::error::ShapeError;
use feotensor::shape;
use feotensor::shape::Shape;
use num::Float;
use std::ops::{Add, Deref, DerefMut, Div, Mul, Sub};

pub struct DeviceMatrix<T: DeviceElem> {
    tensor: DeviceTensor<T>,
}

impl<T: DeviceElem> DeviceMatrix<T> {
    /// matrix.rs:19-23 (like the reference, `new` does not check order == 2; `from_tensor` does)
    pub fn new(shape: &Shape, data: &[T]) -> Result<DeviceMatrix<T>, ShapeError> {
        Ok(DeviceMatrix { tensor: DeviceTensor::new(shape, data)? })
    }
    /// matrix.rs:25-30
    pub fn from_tensor(tensor: DeviceTensor<T>) -> Result<DeviceMatrix<T>, ShapeError> {
        if tensor.shape().order() != 2 {
            return Err(ShapeError::new("Shape must have order of 2"));
        }
        Ok(DeviceMatrix { tensor })
    }
    pub fn into_tensor(self) -> DeviceTensor<T> { self.tensor }
    pub fn fill(shape: &Shape, value: T) -> Result<DeviceMatrix<T>, ShapeError> { Ok(DeviceMatrix { tensor: DeviceTensor::fill(shape, value) }) }
    pub fn zeros(shape: &Shape) -> Result<DeviceMatrix<T>, ShapeError> { Self::fill(shape, T::zero()) }
    pub fn ones(shape: &Shape) -> Result<DeviceMatrix<T>, ShapeError> { Self::fill(shape, T::one()) }
    /// matrix.rs:36-42: panics (through an out-of-bounds `set(..).unwrap()`) when rows > cols
    pub fn eye(shape: &Shape) -> Result<DeviceMatrix<T>, ShapeError> {
        assert!(shape.order() == 2 && shape[0] <= shape[1], "out of bounds for dimension 1");
        let m = Self::zeros(shape)?;
        m.tensor.data.ctx.check(unsafe { ffi::feo_eye(m.tensor.data.ctx.raw, T::DTYPE, m.tensor.data.ptr, shape[0], shape[1]) });
        Ok(m)
    }

    // matrix.rs:50-73: reductions return a DynamicVector through from_tensor(..).unwrap()
    pub fn sum(&self, axes: Axes) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.sum(axes)).unwrap() }
    pub fn mean(&self, axes: Axes) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.mean(axes)).unwrap() }
    pub fn var(&self, axes: Axes) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.var(axes)).unwrap() }
    pub fn max(&self, axes: Axes) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.max(axes)).unwrap() }
    pub fn min(&self, axes: Axes) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.min(axes)).unwrap() }

    /// matrix.rs:76-90.  f32 runs on the tcgen05 tensor cores (3xTF32) when the product is large enough, else SIMT.
    pub fn matmul(&self, rhs: &Self) -> DeviceMatrix<T> {
        assert_eq!(self.shape()[1], rhs.shape()[0]);
        let (m, k, n) = (self.shape()[0], self.shape()[1], rhs.shape()[1]);
        let out = DeviceTensor::uninit(&self.tensor.data.ctx, &shape![m, n].unwrap());
        self.tensor.data.ctx.check(unsafe {
            ffi::feo_matmul(self.tensor.data.ctx.raw, ffi::FEO_MATMUL_AUTO, T::DTYPE, out.data.ptr, self.tensor.data.ptr, rhs.tensor.data.ptr, m, k, n)
        });
        DeviceMatrix::from_tensor(out).unwrap()
    }
    /// matrix.rs:92-103
    pub fn vecmul(&self, rhs: &DeviceVector<T>) -> DeviceVector<T> {
        assert_eq!(self.shape()[1], rhs.shape()[0]);
        let (m, k) = (self.shape()[0], self.shape()[1]);
        let out = DeviceTensor::uninit(&self.tensor.data.ctx, &shape![m].unwrap());
        self.tensor.data.ctx.check(unsafe {
            ffi::feo_matvec(self.tensor.data.ctx.raw, T::DTYPE, out.data.ptr, self.tensor.data.ptr, rhs.data.ptr, m, k)
        });
        DeviceVector::from_tensor(out).unwrap()
    }
}

impl<T: DeviceElem + Float> DeviceMatrix<T> {
    /// matrix.rs:106-110
    pub fn pow(&self, power: T) -> DeviceMatrix<T> { DeviceMatrix::from_tensor(self.tensor.pow(power)).unwrap() }
}

// matrix.rs:113-214
macro_rules! matrix_ops {
    ($tr:ident, $f:ident, $code:expr) => {
        impl<T: DeviceElem> $tr<T> for DeviceMatrix<T> {
            type Output = DeviceMatrix<T>;
            fn $f(self, rhs: T) -> DeviceMatrix<T> { DeviceMatrix::from_tensor(self.tensor.scalar($code, rhs)).unwrap() }
        }
        impl<T: DeviceElem> $tr<DeviceMatrix<T>> for DeviceMatrix<T> {
            type Output = DeviceMatrix<T>;
            fn $f(self, rhs: DeviceMatrix<T>) -> DeviceMatrix<T> { DeviceMatrix::from_tensor(self.tensor.binary($code, rhs.tensor)).unwrap() }
        }
        impl<T: DeviceElem> $tr<DeviceTensor<T>> for DeviceMatrix<T> {
            type Output = DeviceMatrix<T>;
            fn $f(self, rhs: DeviceTensor<T>) -> DeviceMatrix<T> { DeviceMatrix::from_tensor(self.tensor.binary($code, rhs)).unwrap() }
        }
    };
}
matrix_ops!(Add, add, ffi::FEO_ADD);
matrix_ops!(Sub, sub, ffi::FEO_SUB);
matrix_ops!(Mul, mul, ffi::FEO_MUL);
matrix_ops!(Div, div, ffi::FEO_DIV);

// matrix.rs:216-228
impl<T: DeviceElem> Deref for DeviceMatrix<T> {
    type Target = DeviceTensor<T>;
    fn deref(&self) -> &DeviceTensor<T> { &self.tensor }
}
impl<T: DeviceElem> DerefMut for DeviceMatrix<T> {
    fn deref_mut(&mut self) -> &mut DeviceTensor<T> { &mut self.tensor }
}
