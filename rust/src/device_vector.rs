//! `DeviceVector<T>`: the reference's `DynamicVector<T>` newtype (feotensor src/vector.rs:12-230).
use crate::device_matrix::DeviceMatrix;
use crate::device_storage::DeviceElem;
use crate::device_tensor::DeviceTensor;
use crate::ffi;
use feotensor::error::ShapeError;
use feotensor::shape;
use feotensor::shape::Shape;
use num::Float;
use std::ops::{Add, Deref, DerefMut, Div, Mul, Sub};

pub struct DeviceVector<T: DeviceElem> {
    tensor: DeviceTensor<T>,
}

impl<T: DeviceElem> DeviceVector<T> {
    /// vector.rs:18-22
    pub fn new(data: &[T]) -> Result<DeviceVector<T>, ShapeError> {
        Ok(DeviceVector { tensor: DeviceTensor::new(&shape![data.len()].unwrap(), data)? })
    }
    /// vector.rs:24-29
    pub fn from_tensor(tensor: DeviceTensor<T>) -> Result<DeviceVector<T>, ShapeError> {
        if tensor.shape().order() != 1 {
            return Err(ShapeError::new("Shape must have order of 1"));
        }
        Ok(DeviceVector { tensor })
    }
    pub fn into_tensor(self) -> DeviceTensor<T> { self.tensor }
    /// vector.rs:31-37
    pub fn fill(shape: &Shape, value: T) -> Result<DeviceVector<T>, ShapeError> {
        if shape.order() != 1 {
            return Err(ShapeError::new("Shape must have order of 1"));
        }
        Ok(DeviceVector { tensor: DeviceTensor::fill(shape, value) })
    }
    pub fn zeros(shape: &Shape) -> Result<DeviceVector<T>, ShapeError> { Self::fill(shape, T::zero()) }
    pub fn ones(shape: &Shape) -> Result<DeviceVector<T>, ShapeError> { Self::fill(shape, T::one()) }

    // vector.rs:45-68: whole-vector reductions always pass axes = vec![]
    pub fn sum(&self) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.sum(vec![])).unwrap() }
    pub fn mean(&self) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.mean(vec![])).unwrap() }
    pub fn var(&self) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.var(vec![])).unwrap() }
    pub fn max(&self) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.max(vec![])).unwrap() }
    pub fn min(&self) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.min(vec![])).unwrap() }

    /// Dot product (vector.rs:71-78): a one-element vector
    pub fn vecmul(&self, rhs: &DeviceVector<T>) -> DeviceVector<T> {
        assert!(self.shape() == rhs.shape());
        let out = DeviceTensor::uninit(&self.tensor.data.ctx, &shape![1].unwrap());
        self.tensor.data.ctx.check(unsafe {
            ffi::feo_dot(self.tensor.data.ctx.raw, T::DTYPE, out.data.ptr, self.tensor.data.ptr, rhs.tensor.data.ptr, self.size())
        });
        DeviceVector::from_tensor(out).unwrap()
    }
    /// vector . matrix (vector.rs:80-91)
    pub fn matmul(&self, rhs: &DeviceMatrix<T>) -> DeviceVector<T> {
        assert_eq!(self.shape()[0], rhs.shape()[0]);
        let (k, n) = (rhs.shape()[0], rhs.shape()[1]);
        let out = DeviceTensor::uninit(&self.tensor.data.ctx, &shape![n].unwrap());
        self.tensor.data.ctx.check(unsafe {
            ffi::feo_vecmat(self.tensor.data.ctx.raw, T::DTYPE, out.data.ptr, self.tensor.data.ptr, rhs.data.ptr, k, n)
        });
        DeviceVector::from_tensor(out).unwrap()
    }
}

impl<T: DeviceElem + Float> DeviceVector<T> {
    /// vector.rs:94-98
    pub fn pow(&self, power: T) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.pow(power)).unwrap() }
}

// vector.rs:101-202
macro_rules! vector_ops {
    ($tr:ident, $f:ident, $code:expr) => {
        impl<T: DeviceElem> $tr<T> for DeviceVector<T> {
            type Output = DeviceVector<T>;
            fn $f(self, rhs: T) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.scalar($code, rhs)).unwrap() }
        }
        impl<T: DeviceElem> $tr<DeviceVector<T>> for DeviceVector<T> {
            type Output = DeviceVector<T>;
            fn $f(self, rhs: DeviceVector<T>) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.binary($code, rhs.tensor)).unwrap() }
        }
        impl<T: DeviceElem> $tr<DeviceTensor<T>> for DeviceVector<T> {
            type Output = DeviceVector<T>;
            fn $f(self, rhs: DeviceTensor<T>) -> DeviceVector<T> { DeviceVector::from_tensor(self.tensor.binary($code, rhs)).unwrap() }
        }
    };
}
vector_ops!(Add, add, ffi::FEO_ADD);
vector_ops!(Sub, sub, ffi::FEO_SUB);
vector_ops!(Mul, mul, ffi::FEO_MUL);
vector_ops!(Div, div, ffi::FEO_DIV);

// vector.rs:204-216
impl<T: DeviceElem> Deref for DeviceVector<T> {
    type Target = DeviceTensor<T>;
    fn deref(&self) -> &DeviceTensor<T> { &self.tensor }
}
impl<T: DeviceElem> DerefMut for DeviceVector<T> {
    fn deref_mut(&mut self) -> &mut DeviceTensor<T> { &mut self.tensor }
}
