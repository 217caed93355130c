// feotensor.hpp -- header-only C++17 host API over the C ABI (feotensor_b200.h), mirroring the reference's public
// interface for the hot path: same type names, method names, argument meaning and error behaviour.
//
//   reference (Rust, src/)                         here
//   Shape, shape!            shape.rs:6-66          feo::Shape
//   Coordinate, coord!       coordinate.rs:6-73     feo::Coordinate
//   Axes                     axes.rs:1              feo::Axes
//   IndexIterator            iter.rs:6-48           feo::IndexIterator
//   ShapeError               error.rs:3-20          feo::ShapeError      (Err(ShapeError) -> throw)
//   panics (assert!, unwrap)                        feo::Panic
//   DynamicStorage<T>        storage.rs:7-100       feo::DeviceStorage<T> (device resident, same row-major flatten)
//   DynamicTensor/Matrix/Vector  tensor.rs, matrix.rs, vector.rs   feo::Tensor<T>, feo::Matrix<T>, feo::Vector<T>
//
// Operators take their operands by value like the reference (`fn add(self, rhs)`): pass temporaries or std::move; the
// device buffers are released (stream-ordered) when the consumed operands go out of scope.  Reductions, prod, matmul
// and vecmul borrow (`&self`).  There is no CPU fallback: constructing a Context without a CUDA device throws.
#pragma once

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "feotensor_b200.h"

namespace feo {

struct ShapeError : std::runtime_error {  // error.rs:3-20, Display "ShapeError: {reason}"
    std::string reason;
    explicit ShapeError(const std::string &r) : std::runtime_error("ShapeError: " + r), reason(r) {}
};
struct Panic : std::logic_error {  // where the reference panics
    explicit Panic(const std::string &m) : std::logic_error(m) {}
};
struct BackendError : std::runtime_error {
    explicit BackendError(const std::string &m) : std::runtime_error(m) {}
};

using Axes = std::vector<size_t>;  // axes.rs:1

template <class T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = FEO_F32; };
template <> struct dtype_of<double> { static constexpr int value = FEO_F64; };
template <> struct dtype_of<int32_t> { static constexpr int value = FEO_I32; };
template <> struct dtype_of<int64_t> { static constexpr int value = FEO_I64; };
template <> struct dtype_of<uint32_t> { static constexpr int value = FEO_U32; };

class Context {
public:
    explicit Context(int device = 0) {
        if (feo_init(device, &raw_) != FEO_OK) throw BackendError("feo_init failed: no usable CUDA device (feotensor_b200 has no CPU fallback)");
    }
    ~Context() { feo_shutdown(raw_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    feo_ctx *raw() const { return raw_; }
    void check(int rc) const {
        if (rc == FEO_OK) return;
        const std::string msg = feo_last_error(raw_);
        if (rc == FEO_ERR_SHAPE) throw ShapeError(msg);
        if (rc == FEO_ERR_ARITH) throw Panic(msg);  // integer x/0 or MIN/-1: the reference panics (here: at the next sync point)
        throw BackendError(msg);
    }
    void sync() const { check(feo_sync(raw_)); }
    static std::shared_ptr<Context> &current() {
        static std::shared_ptr<Context> ctx;
        if (!ctx) ctx = std::make_shared<Context>(0);
        return ctx;
    }

private:
    feo_ctx *raw_ = nullptr;
};

class Shape {  // shape.rs:6-56
public:
    explicit Shape(std::vector<size_t> dims) : dims_(std::move(dims)) {
        if (feo_shape_check((int)dims_.size(), dims_.data()) != FEO_OK) throw ShapeError("Dimension cannot be zero");  // shape.rs:13-15
    }
    Shape(std::initializer_list<size_t> dims) : Shape(std::vector<size_t>(dims)) {}  // the shape! macro
    size_t size() const { return feo_shape_size((int)dims_.size(), dims_.data()); }
    size_t order() const { return dims_.size(); }
    Shape stack(const Shape &rhs) const {
        std::vector<size_t> d = dims_;
        d.insert(d.end(), rhs.dims_.begin(), rhs.dims_.end());
        return Shape(d);
    }
    size_t operator[](size_t i) const { return dims_.at(i); }
    const std::vector<size_t> &dims() const { return dims_; }
    bool operator==(const Shape &o) const { return dims_ == o.dims_; }
    bool operator!=(const Shape &o) const { return !(*this == o); }
    std::string to_string() const { return join(dims_); }  // Display "(2, 3, 4)"
    static std::string join(const std::vector<size_t> &v) {
        std::string s = "(";
        for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + std::to_string(v[i]);
        return s + ")";
    }

private:
    std::vector<size_t> dims_;
};

class Coordinate {  // coordinate.rs:6-56
public:
    explicit Coordinate(std::vector<size_t> indices) : indices_(std::move(indices)) {
        if (indices_.empty()) throw ShapeError("Coordinate cannot be empty");
    }
    Coordinate(std::initializer_list<size_t> i) : Coordinate(std::vector<size_t>(i)) {}  // the coord! macro
    static Coordinate repeat(size_t index, size_t count) { return Coordinate(std::vector<size_t>(count, index)); }
    size_t order() const { return indices_.size(); }
    Coordinate insert(size_t index, size_t axis) const {  // coordinate.rs:27-33: inserts the VALUE `axis` at `index`
        if (index > indices_.size()) throw Panic("insertion index out of bounds");
        std::vector<size_t> n = indices_;
        n.insert(n.begin() + (std::ptrdiff_t)index, axis);
        return Coordinate(n);
    }
    size_t operator[](size_t i) const { return indices_.at(i); }
    size_t &operator[](size_t i) { return indices_.at(i); }
    const std::vector<size_t> &indices() const { return indices_; }
    bool operator==(const Coordinate &o) const { return indices_ == o.indices_; }
    std::string to_string() const { return Shape::join(indices_); }

private:
    std::vector<size_t> indices_;
};

class IndexIterator {  // iter.rs:6-48: row-major odometer, nothing for order-0 shapes
public:
    explicit IndexIterator(const Shape &shape) : shape_(shape), current_(std::max<size_t>(shape.order(), 1), 0) {}
    bool next(Coordinate &out) {
        if (done_ || shape_.order() == 0) return false;
        out = Coordinate(current_);
        for (size_t i = shape_.order(); i-- > 0;) {
            if (current_[i] + 1 < shape_[i]) { current_[i] += 1; break; }
            current_[i] = 0;
            if (i == 0) done_ = true;
        }
        return true;
    }

private:
    Shape shape_;
    std::vector<size_t> current_;
    bool done_ = false;
};

template <class T> class DeviceStorage {  // beside DynamicStorage<T> (storage.rs:7-100)
public:
    DeviceStorage(std::shared_ptr<Context> ctx, size_t n) : ctx_(std::move(ctx)), n_(n) { ctx_->check(feo_malloc(ctx_->raw(), n * sizeof(T), &ptr_)); }
    DeviceStorage(std::shared_ptr<Context> ctx, const std::vector<T> &data) : DeviceStorage(std::move(ctx), data.size()) {
        ctx_->check(feo_upload(ctx_->raw(), ptr_, data.data(), data.size() * sizeof(T)));
        ctx_->sync();  // `data` may be a temporary: the pageable copy must have left the host buffer
    }
    DeviceStorage(DeviceStorage &&o) noexcept : ctx_(std::move(o.ctx_)), ptr_(o.ptr_), n_(o.n_) { o.ptr_ = nullptr; }
    DeviceStorage &operator=(DeviceStorage &&o) noexcept {
        if (this != &o) { release(); ctx_ = std::move(o.ctx_); ptr_ = o.ptr_; n_ = o.n_; o.ptr_ = nullptr; }
        return *this;
    }
    DeviceStorage(const DeviceStorage &) = delete;  // the reference's storage is not Clone either (storage.rs:7)
    ~DeviceStorage() { release(); }

    size_t flatten(const Coordinate &c, const Shape &s) const {  // storage.rs:18-38
        size_t idx = 0;
        int detail = 0;
        if (feo_flatten((int)c.order(), c.indices().data(), (int)s.order(), s.dims().data(), &idx, &detail) != FEO_OK) {
            if (detail < 0) throw ShapeError("incorrect order (" + std::to_string(c.order()) + " vs " + std::to_string(s.order()) + ").");
            throw ShapeError("out of bounds for dimension " + std::to_string(detail));
        }
        return idx;
    }
    size_t size() const { return n_; }
    T at(size_t i) const {
        if (i >= n_) throw Panic("index out of bounds");
        T v;
        ctx_->check(feo_download(ctx_->raw(), &v, static_cast<const char *>(ptr_) + i * sizeof(T), sizeof(T)));
        return v;
    }
    void set(size_t i, T v) {
        if (i >= n_) throw Panic("index out of bounds");
        ctx_->check(feo_upload(ctx_->raw(), static_cast<char *>(ptr_) + i * sizeof(T), &v, sizeof(T)));
        ctx_->sync();
    }
    std::vector<T> to_vec() const {
        std::vector<T> v(n_);
        ctx_->check(feo_download(ctx_->raw(), v.data(), ptr_, n_ * sizeof(T)));
        return v;
    }
    void *ptr() const { return ptr_; }
    const std::shared_ptr<Context> &ctx() const { return ctx_; }

private:
    void release() {
        if (ptr_ && ctx_) feo_free(ctx_->raw(), ptr_);
        ptr_ = nullptr;
    }
    std::shared_ptr<Context> ctx_;
    void *ptr_ = nullptr;
    size_t n_ = 0;
};

template <class T> class Vector;
template <class T> class Matrix;

template <class T> class Tensor {  // DynamicTensor<T> (tensor.rs:15-19)
public:
    static constexpr int DT = dtype_of<T>::value;

    Tensor(const Shape &shape, const std::vector<T> &data) : shape_(shape), data_(make_storage(shape, data)) {}  // Tensor::new (tensor.rs:22-30)
    static Tensor fill(const Shape &shape, T value) {  // tensor.rs:32-38
        Tensor t(shape, DeviceStorage<T>(Context::current(), shape.size()));
        t.ctx().check(feo_fill(t.ctx().raw(), DT, t.ptr(), &value, shape.size()));
        return t;
    }
    static Tensor zeros(const Shape &shape) { return fill(shape, T(0)); }
    static Tensor ones(const Shape &shape) { return fill(shape, T(1)); }

    const Shape &shape() const { return shape_; }
    size_t size() const { return shape_.size(); }
    T get(const Coordinate &c) const { return data_.at(data_.flatten(c, shape_)); }           // tensor.rs:54-56
    void set(const Coordinate &c, T value) { data_.set(data_.flatten(c, shape_), value); }    // tensor.rs:63-67
    std::vector<T> to_vec() const { return data_.to_vec(); }
    const DeviceStorage<T> &data() const { return data_; }
    void *ptr() const { return data_.ptr(); }
    const Context &ctx() const { return *data_.ctx(); }

    // reductions (tensor.rs:70-307)
    Tensor sum(const Axes &axes) const { return reduce(FEO_SUM, axes); }
    Tensor mean(const Axes &axes) const { return reduce(FEO_MEAN, axes); }
    Tensor var(const Axes &axes) const { return reduce(FEO_VAR, axes); }
    Tensor max(const Axes &axes) const { return reduce(FEO_MAX, axes); }
    Tensor min(const Axes &axes) const { return reduce(FEO_MIN, axes); }

    Tensor pow(T power) const {  // tensor.rs:325-333 (Float element types only)
        static_assert(std::is_floating_point<T>::value, "pow needs a Float element type (tensor.rs:325)");
        Tensor out(shape_, DeviceStorage<T>(data_.ctx(), size()));
        ctx().check(feo_pow(ctx().raw(), DT, out.ptr(), ptr(), &power, size()));
        return out;
    }
    Tensor prod(const Tensor &other) const {  // tensor.rs:311-322
        Tensor out(shape_.stack(other.shape_), DeviceStorage<T>(data_.ctx(), size() * other.size()));
        ctx().check(feo_outer(ctx().raw(), DT, out.ptr(), ptr(), size(), other.ptr(), other.size()));
        return out;
    }
    // broadcasting elementwise: NEW surface (the reference asserts equal shapes)
    Tensor broadcast(int op, const Tensor &rhs) const {
        const auto &a = shape_.dims(), &b = rhs.shape_.dims();
        if (a.size() != b.size()) throw ShapeError("incorrect order (" + std::to_string(a.size()) + " vs " + std::to_string(b.size()) + ").");
        std::vector<size_t> od(a.size()), sa(a.size()), sb(a.size());
        for (size_t d = 0; d < a.size(); ++d) {
            if (a[d] != b[d] && a[d] != 1 && b[d] != 1) throw ShapeError("shapes are not broadcastable");
            od[d] = a[d] > b[d] ? a[d] : b[d];
        }
        size_t xa = 1, xb = 1;
        for (size_t d = a.size(); d-- > 0;) {
            sa[d] = (a[d] == od[d] && a[d] != 1) ? xa : 0;
            sb[d] = (b[d] == od[d] && b[d] != 1) ? xb : 0;
            xa *= a[d];
            xb *= b[d];
        }
        Shape os(od);
        Tensor out(os, DeviceStorage<T>(data_.ctx(), os.size()));
        ctx().check(feo_ewise_broadcast(ctx().raw(), op, DT, out.ptr(), ptr(), rhs.ptr(), (int)od.size(), od.data(), sa.data(), sb.data()));
        return out;
    }

    // engine of the operator impls (tensor.rs:336-489)
    static Tensor binary(int op, const Tensor &a, const Tensor &b) {
        if (a.shape_ != b.shape_) throw Panic("assertion failed: self.shape == rhs.shape");  // tensor.rs:366
        Tensor out(a.shape_, DeviceStorage<T>(a.data_.ctx(), a.size()));
        a.ctx().check(feo_ewise(a.ctx().raw(), op, DT, out.ptr(), a.ptr(), b.ptr(), a.size()));
        return out;
    }
    static Tensor scalar(int op, const Tensor &a, T s) {
        Tensor out(a.shape_, DeviceStorage<T>(a.data_.ctx(), a.size()));
        a.ctx().check(feo_ewise_scalar(a.ctx().raw(), op, DT, out.ptr(), a.ptr(), &s, a.size()));
        return out;
    }
    Tensor(const Shape &shape, DeviceStorage<T> &&storage) : shape_(shape), data_(std::move(storage)) {}
    Tensor(Tensor &&) noexcept = default;
    Tensor &operator=(Tensor &&) noexcept = default;

private:
    static DeviceStorage<T> make_storage(const Shape &shape, const std::vector<T> &data) {
        if (data.size() != shape.size()) throw ShapeError("Data length does not match shape size");  // tensor.rs:23-25
        return DeviceStorage<T>(Context::current(), data);
    }
    Tensor reduce(int kind, const Axes &axes) const {
        int out_rank = 0;
        std::vector<size_t> od(shape_.order() ? shape_.order() : 1);
        if (feo_reduce_shape((int)shape_.order(), shape_.dims().data(), (int)axes.size(), axes.data(), &out_rank, od.data()) != FEO_OK)
            throw Panic("axes must be strictly ascending, unique and < order (the reference panics)");
        od.resize((size_t)out_rank);
        Shape os(od);
        Tensor out(os, DeviceStorage<T>(data_.ctx(), os.size()));
        ctx().check(feo_reduce(ctx().raw(), kind, DT, out.ptr(), ptr(), (int)shape_.order(), shape_.dims().data(), (int)axes.size(), axes.data()));
        return out;
    }
    Shape shape_;
    DeviceStorage<T> data_;
};

// std::ops impls on Tensor: operands by value (tensor.rs:336-489)
template <class T> Tensor<T> operator+(Tensor<T> a, Tensor<T> b) { return Tensor<T>::binary(FEO_ADD, a, b); }
template <class T> Tensor<T> operator-(Tensor<T> a, Tensor<T> b) { return Tensor<T>::binary(FEO_SUB, a, b); }
template <class T> Tensor<T> operator*(Tensor<T> a, Tensor<T> b) { return Tensor<T>::binary(FEO_MUL, a, b); }
template <class T> Tensor<T> operator/(Tensor<T> a, Tensor<T> b) { return Tensor<T>::binary(FEO_DIV, a, b); }
template <class T> Tensor<T> operator+(Tensor<T> a, T s) { return Tensor<T>::scalar(FEO_ADD, a, s); }
template <class T> Tensor<T> operator-(Tensor<T> a, T s) { return Tensor<T>::scalar(FEO_SUB, a, s); }
template <class T> Tensor<T> operator*(Tensor<T> a, T s) { return Tensor<T>::scalar(FEO_MUL, a, s); }
template <class T> Tensor<T> operator/(Tensor<T> a, T s) { return Tensor<T>::scalar(FEO_DIV, a, s); }

template <class T> class Vector {  // DynamicVector<T> (vector.rs:12-15)
public:
    explicit Vector(const std::vector<T> &data) : tensor_(Shape({data.size()}), data) {}  // vector.rs:18-22
    static Vector from_tensor(Tensor<T> t) {                                                // vector.rs:24-29
        if (t.shape().order() != 1) throw ShapeError("Shape must have order of 1");
        return Vector(std::move(t), 0);
    }
    static Vector fill(const Shape &s, T v) {                                                // vector.rs:31-37
        if (s.order() != 1) throw ShapeError("Shape must have order of 1");
        return Vector(Tensor<T>::fill(s, v), 0);
    }
    static Vector zeros(const Shape &s) { return fill(s, T(0)); }
    static Vector ones(const Shape &s) { return fill(s, T(1)); }
    // Deref<Target = DynamicTensor> (vector.rs:204-216)
    const Tensor<T> &tensor() const { return tensor_; }
    Tensor<T> into_tensor() && { return std::move(tensor_); }
    const Shape &shape() const { return tensor_.shape(); }
    size_t size() const { return tensor_.size(); }
    T get(const Coordinate &c) const { return tensor_.get(c); }
    void set(const Coordinate &c, T v) { tensor_.set(c, v); }
    Tensor<T> prod(const Vector &o) const { return tensor_.prod(o.tensor_); }
    Vector pow(T power) const { return from_tensor(tensor_.pow(power)); }  // vector.rs:94-98
    std::vector<T> to_vec() const { return tensor_.to_vec(); }
    T operator[](size_t i) const {  // Index<usize> (vector.rs:218-224): unwrap -> panic
        try { return tensor_.get(Coordinate({i})); } catch (const ShapeError &e) { throw Panic(e.what()); }
    }
    // whole-vector reductions always pass axes = [] (vector.rs:45-68)
    Vector sum() const { return from_tensor(tensor_.sum({})); }
    Vector mean() const { return from_tensor(tensor_.mean({})); }
    Vector var() const { return from_tensor(tensor_.var({})); }
    Vector max() const { return from_tensor(tensor_.max({})); }
    Vector min() const { return from_tensor(tensor_.min({})); }
    Vector vecmul(const Vector &rhs) const {  // dot product (vector.rs:71-78)
        if (shape() != rhs.shape()) throw Panic("assertion failed: self.shape() == rhs.shape()");
        Tensor<T> out(Shape({1}), DeviceStorage<T>(tensor_.data().ctx(), 1));
        tensor_.ctx().check(feo_dot(tensor_.ctx().raw(), Tensor<T>::DT, out.ptr(), tensor_.ptr(), rhs.tensor_.ptr(), size()));
        return from_tensor(std::move(out));
    }
    Vector matmul(const Matrix<T> &rhs) const;  // vector.rs:80-91

private:
    Vector(Tensor<T> &&t, int) : tensor_(std::move(t)) {}
    Tensor<T> tensor_;
    friend class Matrix<T>;
};

template <class T> class Matrix {  // DynamicMatrix<T> (matrix.rs:13-16)
public:
    Matrix(const Shape &shape, const std::vector<T> &data) : tensor_(shape, data) {}  // matrix.rs:19-23 (no order check here)
    static Matrix from_tensor(Tensor<T> t) {                                          // matrix.rs:25-30
        if (t.shape().order() != 2) throw ShapeError("Shape must have order of 2");
        return Matrix(std::move(t), 0);
    }
    static Matrix fill(const Shape &s, T v) { return Matrix(Tensor<T>::fill(s, v), 0); }
    static Matrix zeros(const Shape &s) { return fill(s, T(0)); }
    static Matrix ones(const Shape &s) { return fill(s, T(1)); }
    static Matrix eye(const Shape &s) {  // matrix.rs:36-42
        if (s.order() != 2 || s[0] > s[1]) throw Panic("called `Result::unwrap()` on an `Err` value: ShapeError: out of bounds");
        Matrix m(Tensor<T>(s, DeviceStorage<T>(Context::current(), s.size())), 0);
        m.tensor_.ctx().check(feo_eye(m.tensor_.ctx().raw(), Tensor<T>::DT, m.tensor_.ptr(), s[0], s[1]));
        return m;
    }
    const Tensor<T> &tensor() const { return tensor_; }
    Tensor<T> into_tensor() && { return std::move(tensor_); }
    const Shape &shape() const { return tensor_.shape(); }
    size_t size() const { return tensor_.size(); }
    T get(const Coordinate &c) const { return tensor_.get(c); }
    void set(const Coordinate &c, T v) { tensor_.set(c, v); }
    std::vector<T> to_vec() const { return tensor_.to_vec(); }
    Matrix pow(T power) const { return from_tensor(tensor_.pow(power)); }  // matrix.rs:106-110
    T operator[](const Coordinate &c) const {  // Index<Coordinate> (matrix.rs:230-236)
        try { return tensor_.get(c); } catch (const ShapeError &e) { throw Panic(e.what()); }
    }
    // reductions return DynamicVector through from_tensor(..).unwrap() (matrix.rs:50-73)
    Vector<T> sum(const Axes &a) const { return unwrap(tensor_.sum(a)); }
    Vector<T> mean(const Axes &a) const { return unwrap(tensor_.mean(a)); }
    Vector<T> var(const Axes &a) const { return unwrap(tensor_.var(a)); }
    Vector<T> max(const Axes &a) const { return unwrap(tensor_.max(a)); }
    Vector<T> min(const Axes &a) const { return unwrap(tensor_.min(a)); }
    Matrix matmul(const Matrix &rhs, int algo = FEO_MATMUL_AUTO) const {  // matrix.rs:76-90
        if (shape()[1] != rhs.shape()[0]) throw Panic("assertion `left == right` failed");
        const size_t M = shape()[0], K = shape()[1], N = rhs.shape()[1];
        Tensor<T> out(Shape({M, N}), DeviceStorage<T>(tensor_.data().ctx(), M * N));
        tensor_.ctx().check(feo_matmul(tensor_.ctx().raw(), algo, Tensor<T>::DT, out.ptr(), tensor_.ptr(), rhs.tensor_.ptr(), M, K, N));
        return from_tensor(std::move(out));
    }
    Vector<T> vecmul(const Vector<T> &rhs) const {  // matrix.rs:92-103
        if (shape()[1] != rhs.shape()[0]) throw Panic("assertion `left == right` failed");
        const size_t M = shape()[0], K = shape()[1];
        Tensor<T> out(Shape({M}), DeviceStorage<T>(tensor_.data().ctx(), M));
        tensor_.ctx().check(feo_matvec(tensor_.ctx().raw(), Tensor<T>::DT, out.ptr(), tensor_.ptr(), rhs.tensor().ptr(), M, K));
        return Vector<T>::from_tensor(std::move(out));
    }

private:
    Matrix(Tensor<T> &&t, int) : tensor_(std::move(t)) {}
    static Vector<T> unwrap(Tensor<T> t) {
        try { return Vector<T>::from_tensor(std::move(t)); } catch (const ShapeError &e) { throw Panic(std::string("called `Result::unwrap()` on an `Err` value: ") + e.what()); }
    }
    Tensor<T> tensor_;
    friend class Vector<T>;
};

template <class T> Vector<T> Vector<T>::matmul(const Matrix<T> &rhs) const {
    if (shape()[0] != rhs.shape()[0]) throw Panic("assertion `left == right` failed");
    const size_t K = rhs.shape()[0], N = rhs.shape()[1];
    Tensor<T> out(Shape({N}), DeviceStorage<T>(tensor_.data().ctx(), N));
    tensor_.ctx().check(feo_vecmat(tensor_.ctx().raw(), Tensor<T>::DT, out.ptr(), tensor_.ptr(), rhs.tensor().ptr(), K, N));
    return from_tensor(std::move(out));
}

// Newtype operator impls (matrix.rs:113-214, vector.rs:101-202) and the mixed Tensor forms (tensor.rs:375-505):
// W (op) {T, W, Tensor} -> W;   Tensor (op) W -> W.
#define FEO_WRAPPER_OPS(W, OPSYM, OPCODE)                                                                                      \
    template <class T> W<T> operator OPSYM(W<T> a, W<T> b) { return W<T>::from_tensor(Tensor<T>::binary(OPCODE, a.tensor(), b.tensor())); } \
    template <class T> W<T> operator OPSYM(W<T> a, Tensor<T> b) { return W<T>::from_tensor(Tensor<T>::binary(OPCODE, a.tensor(), b)); }     \
    template <class T> W<T> operator OPSYM(Tensor<T> a, W<T> b) { return W<T>::from_tensor(Tensor<T>::binary(OPCODE, a, b.tensor())); }     \
    template <class T> W<T> operator OPSYM(W<T> a, T s) { return W<T>::from_tensor(Tensor<T>::scalar(OPCODE, a.tensor(), s)); }
FEO_WRAPPER_OPS(Vector, +, FEO_ADD) FEO_WRAPPER_OPS(Vector, -, FEO_SUB) FEO_WRAPPER_OPS(Vector, *, FEO_MUL) FEO_WRAPPER_OPS(Vector, /, FEO_DIV)
FEO_WRAPPER_OPS(Matrix, +, FEO_ADD) FEO_WRAPPER_OPS(Matrix, -, FEO_SUB) FEO_WRAPPER_OPS(Matrix, *, FEO_MUL) FEO_WRAPPER_OPS(Matrix, /, FEO_DIV)
#undef FEO_WRAPPER_OPS

// ---- cross-GPU epilogues over peer memory (feotensor_b200.h, "cross-GPU epilogues"; one process per GPU) ---------------------
// The host program carries the 64-byte handles between the ranks with whatever rendezvous it has (MPI, a file, a socket).
using PeerHandle = std::array<unsigned char, FEO_PEER_HANDLE_BYTES>;

class PeerWindow {  // peer-mappable device memory owned by this rank
public:
    PeerWindow(std::shared_ptr<Context> ctx, size_t bytes) : ctx_(std::move(ctx)), bytes_(bytes) { ctx_->check(feo_peer_alloc(ctx_->raw(), bytes, &ptr_)); }
    ~PeerWindow() { if (ptr_) feo_peer_free(ctx_->raw(), ptr_); }
    PeerWindow(const PeerWindow &) = delete;
    PeerWindow &operator=(const PeerWindow &) = delete;
    PeerHandle export_handle() const {
        PeerHandle h{};
        ctx_->check(feo_peer_export(ctx_->raw(), ptr_, h.data()));
        return h;
    }
    void *ptr() const { return ptr_; }
    size_t bytes() const { return bytes_; }

private:
    std::shared_ptr<Context> ctx_;
    void *ptr_ = nullptr;
    size_t bytes_;
};

class PeerComm {  // the `world` windows of one node as seen from this rank; every rank issues the same sequence of calls
public:
    // handles[r] = rank r's PeerWindow::export_handle(); `own` is this rank's window.  The caller synchronises the ranks
    // (all windows created and mapped) before the first collective call.
    PeerComm(std::shared_ptr<Context> ctx, int rank, std::shared_ptr<PeerWindow> own, const std::vector<PeerHandle> &handles)
        : ctx_(std::move(ctx)), own_(std::move(own)), rank_(rank), world_((int)handles.size()) {
        std::vector<void *> windows(handles.size());
        for (size_t r = 0; r < handles.size(); ++r) {
            if ((int)r == rank) { windows[r] = own_->ptr(); continue; }
            ctx_->check(feo_peer_open(ctx_->raw(), handles[r].data(), &windows[r]));
            opened_.push_back(windows[r]);
        }
        ctx_->check(feo_comm_create(ctx_->raw(), rank, world_, windows.data(), own_->bytes(), &raw_));
    }
    ~PeerComm() {
        feo_sync(ctx_->raw());
        feo_comm_destroy(raw_);
        for (void *p : opened_) feo_peer_close(ctx_->raw(), p);
    }
    PeerComm(const PeerComm &) = delete;
    PeerComm &operator=(const PeerComm &) = delete;
    int rank() const { return rank_; }
    int world() const { return world_; }
    void barrier() const { ctx_->check(feo_comm_barrier(raw_)); }

    // Tensor::{sum,mean,var,max,min} (tensor.rs:70-307) of the GLOBAL tensor whose block of leading rows is `local`; `axes`
    // contains 0 or is empty (otherwise no exchange is needed: call the local method).  Replicated, bit-identical result.
    template <class T> Tensor<T> reduce(int kind, const Tensor<T> &local, const Shape &global, const Axes &axes) const {
        int out_rank = 0;
        std::vector<size_t> od(global.order() ? global.order() : 1);
        if (feo_reduce_shape((int)global.order(), global.dims().data(), (int)axes.size(), axes.data(), &out_rank, od.data()) != FEO_OK)
            throw Panic("axes must be strictly ascending, unique and < order (the reference panics)");
        od.resize((size_t)out_rank);
        T divisor{};  // the reference counts it in T over the GLOBAL shape (tensor.rs:112-130)
        feo_reduce_divisor(Tensor<T>::DT, (int)global.order(), global.dims().data(), (int)axes.size(), axes.data(), &divisor);
        Axes local_axes = axes;
        if (local_axes.empty())
            for (size_t d = 0; d < local.shape().order(); ++d) local_axes.push_back(d);  // axes == [] reduces everything (tensor.rs:84)
        Shape os(od);
        Tensor<T> out(os, DeviceStorage<T>(local.data().ctx(), os.size()));
        ctx_->check(feo_reduce_sharded(raw_, kind, Tensor<T>::DT, out.ptr(), local.ptr(), (int)local.shape().order(), local.shape().dims().data(),
                                       (int)local_axes.size(), local_axes.data(), &divisor));
        return out;
    }
    // DynamicVector::vecmul (vector.rs:71-78) on two identically sharded vectors: a replicated one-element result.
    template <class T> Tensor<T> dot(const Tensor<T> &a, const Tensor<T> &b) const {
        if (a.shape() != b.shape()) throw Panic("assertion failed: self.shape() == rhs.shape()");
        Tensor<T> out(Shape({1}), DeviceStorage<T>(a.data().ctx(), 1));
        ctx_->check(feo_dot_sharded(raw_, Tensor<T>::DT, out.ptr(), a.ptr(), b.ptr(), 1, a.size()));
        return out;
    }
    // DynamicMatrix::matmul (matrix.rs:76-90), A row-sharded (a_local = this rank's [M, K] rows) and B row-sharded:
    // b_panels[r] = rank r's [K / world, N] panel in peer-mapped memory; returns this rank's [M, N] rows.
    template <class T> Tensor<T> matmul_sharded_b(const Tensor<T> &a_local, const std::vector<const void *> &b_panels, size_t K, size_t N) const {
        const size_t M = a_local.shape()[0];
        if (a_local.shape()[1] != K) throw Panic("assertion `left == right` failed");
        Tensor<T> out(Shape({M, N}), DeviceStorage<T>(a_local.data().ctx(), M * N));
        ctx_->check(feo_matmul_sharded_b(raw_, FEO_MATMUL_AUTO, Tensor<T>::DT, out.ptr(), a_local.ptr(), b_panels.data(), M, K, N));
        return out;
    }

private:
    std::shared_ptr<Context> ctx_;
    std::shared_ptr<PeerWindow> own_;
    feo_comm *raw_ = nullptr;
    std::vector<void *> opened_;
    int rank_, world_;
};

// ---- the NCCL route of the same exchanges, issued inside the C ABI (csrc/nccl_route.cu) ------------------------------------
// One process (or thread) per GPU: rank 0 creates the id, the host program's own rendezvous (MPI_Bcast, a file, a socket)
// carries its 128 bytes to the other ranks, every rank constructs an NcclComm.  libnccl.so.2 is dlopen'ed by the library.
using NcclId = std::array<unsigned char, FEO_NCCL_UNIQUE_ID_BYTES>;

class NcclComm {
public:
    static NcclId unique_id() {
        NcclId id{};
        if (feo_nccl_unique_id(id.data()) != FEO_OK) throw BackendError("feo_nccl_unique_id failed: is libnccl.so.2 on the loader path (FEO_NCCL_LIB)?");
        return id;
    }
    NcclComm(std::shared_ptr<Context> ctx, const NcclId &id, int rank, int world) : ctx_(std::move(ctx)), rank_(rank), world_(world) {
        ctx_->check(feo_nccl_comm_create(ctx_->raw(), id.data(), rank, world, &raw_));
    }
    ~NcclComm() {
        feo_sync(ctx_->raw());
        feo_nccl_comm_destroy(raw_);
    }
    NcclComm(const NcclComm &) = delete;
    NcclComm &operator=(const NcclComm &) = delete;
    int rank() const { return rank_; }
    int world() const { return world_; }

    // Tensor::{sum,mean,var,max,min} (tensor.rs:70-307) of the GLOBAL tensor whose block of leading rows is `local`: local
    // one-pass partial, ncclAllReduce (var: ncclAllGather + rank-ordered merge), epilogue.  Same contract as PeerComm::reduce.
    template <class T> Tensor<T> reduce(int kind, const Tensor<T> &local, const Shape &global, const Axes &axes) const {
        int out_rank = 0;
        std::vector<size_t> od(global.order() ? global.order() : 1);
        if (feo_reduce_shape((int)global.order(), global.dims().data(), (int)axes.size(), axes.data(), &out_rank, od.data()) != FEO_OK)
            throw Panic("axes must be strictly ascending, unique and < order (the reference panics)");
        od.resize((size_t)out_rank);
        T divisor{};
        feo_reduce_divisor(Tensor<T>::DT, (int)global.order(), global.dims().data(), (int)axes.size(), axes.data(), &divisor);
        Axes local_axes = axes;
        if (local_axes.empty())
            for (size_t d = 0; d < local.shape().order(); ++d) local_axes.push_back(d);
        Shape os(od);
        Tensor<T> out(os, DeviceStorage<T>(local.data().ctx(), os.size()));
        ctx_->check(feo_reduce_sharded_nccl(raw_, kind, Tensor<T>::DT, out.ptr(), local.ptr(), (int)local.shape().order(),
                                            local.shape().dims().data(), (int)local_axes.size(), local_axes.data(), &divisor));
        return out;
    }
    // DynamicVector::vecmul (vector.rs:71-78) on two identically sharded vectors: partial dots, then ncclAllReduce.
    template <class T> Tensor<T> dot(const Tensor<T> &a, const Tensor<T> &b) const {
        if (a.shape() != b.shape()) throw Panic("assertion failed: self.shape() == rhs.shape()");
        Tensor<T> out(Shape({1}), DeviceStorage<T>(a.data().ctx(), 1));
        ctx_->check(feo_dot_sharded_nccl(raw_, Tensor<T>::DT, out.ptr(), a.ptr(), b.ptr(), 1, a.size()));
        return out;
    }
    // DynamicMatrix::matmul (matrix.rs:76-90), A and B row-sharded: ncclAllGather of B (b_local = this rank's [K / world, N]
    // panel), then the local kernel; returns this rank's [M, N] rows.
    template <class T> Tensor<T> matmul_sharded_b(const Tensor<T> &a_local, const Tensor<T> &b_local, size_t K, size_t N) const {
        const size_t M = a_local.shape()[0];
        if (a_local.shape()[1] != K) throw Panic("assertion `left == right` failed");
        Tensor<T> out(Shape({M, N}), DeviceStorage<T>(a_local.data().ctx(), M * N));
        ctx_->check(feo_matmul_sharded_b_nccl(raw_, FEO_MATMUL_AUTO, Tensor<T>::DT, out.ptr(), a_local.ptr(), b_local.ptr(), M, K, N));
        return out;
    }

private:
    std::shared_ptr<Context> ctx_;
    feo_nccl *raw_ = nullptr;
    int rank_, world_;
};

}  // namespace feo
