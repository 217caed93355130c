//! feotensor-b200: a B200 (sm_100a) compute backend for feotensor.
//!
//! `DeviceStorage<T>` sits beside the reference's `DynamicStorage<T>`; `DeviceTensor`, `DeviceMatrix` and `DeviceVector`
//! mirror `DynamicTensor`, `DynamicMatrix` and `DynamicVector` (same methods, same operator impls, same error and panic
//! behaviour) and forward to the C ABI of `include/feotensor_b200.h`.  `Shape`, `Coordinate`, `Axes` and `ShapeError`
//! are the reference's own types.
//!
//! ```ignore
//! use feotensor::shape;
//! use feotensor_b200::DeviceTensor;
//! let a = DeviceTensor::<f32>::new(&shape![64, 128, 256].unwrap(), &host_a).unwrap();
//! let b = DeviceTensor::<f32>::new(&shape![64, 128, 256].unwrap(), &host_b).unwrap();
//! let s = (a + b).sum(vec![0, 2]);          // one elementwise kernel, one reduction; operands are consumed
//! let host: Vec<f32> = s.to_vec();
//! ```
pub mod comm;
pub mod context;
pub mod device_matrix;
pub mod device_storage;
pub mod device_tensor;
pub mod device_vector;
pub mod ffi;

pub use comm::{NcclComm, NcclId, PeerComm, PeerHandle, PeerWindow};
pub use context::Context;
pub use device_matrix::DeviceMatrix;
pub use device_storage::{DeviceElem, DeviceStorage};
pub use device_tensor::DeviceTensor;
pub use device_vector::DeviceVector;

pub type Tensor<T> = DeviceTensor<T>;
pub type Matrix<T> = DeviceMatrix<T>;
pub type Vector<T> = DeviceVector<T>;
