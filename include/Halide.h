/*
 * Halide.h -- minimal stand-in for the Halide types that leak through the RecFilter
 * operator surface (/root/reference/lib/recfilter.h), so that the reference's tests and
 * apps compile unchanged against the B200 engine.  This is NOT Halide: there is no
 * compiler, no JIT and no scheduling; only the value types the programs touch.
 *
 * What the in-scope programs use:
 *   Image<T>      dense host array, dimension 0 contiguous (the reference's layout,
 *                 lib/recfilter.cpp:970-981); (w[,h[,c[,d]]]) constructors, construction from a
 *                 Realization / Buffer, operator()(int...) -> T&, operator()(Expr...) -> Expr,
 *                 width/height/channels/dimensions/min/extent
 *   ImageParam    (Type, dims) + set(Image) + operator()(Expr...)
 *   Buffer / Realization   what RecFilter::realize() returns
 *   Expr          tiny expression tree: variables, constants, + - * / min max clamp, image loads
 *                 and Func calls -- enough to describe "which array is filtered"
 *   Var, Func (operator() and chainable no-op schedule calls), Tuple, Type / type_of<T>(), Target
 */
#ifndef RECFILTER_B200_HALIDE_SHIM_H_
#define RECFILTER_B200_HALIDE_SHIM_H_

#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

class RecFilter;
struct RecFilterContents;

namespace Halide {

// ---------------------------------------------------------------------------------------------
// element types
// ---------------------------------------------------------------------------------------------
struct Type {
    enum Code { Int = 0, UInt = 1, Float = 2 };
    Code code = Float;
    int bits = 32;
    Type() {}
    Type(Code c, int b) : code(c), bits(b) {}
    int bytes() const { return bits / 8; }
    bool is_float() const { return code == Float; }
    bool is_int() const { return code == Int; }
    bool is_uint() const { return code == UInt; }
    bool operator==(const Type& o) const { return code == o.code && bits == o.bits; }
    bool operator!=(const Type& o) const { return !(*this == o); }
};
inline Type Int(int bits)   { return Type(Type::Int, bits); }
inline Type UInt(int bits)  { return Type(Type::UInt, bits); }
inline Type Float(int bits) { return Type(Type::Float, bits); }

template <typename T> inline Type type_of();
template <> inline Type type_of<float>()    { return Float(32); }
template <> inline Type type_of<double>()   { return Float(64); }
template <> inline Type type_of<int8_t>()   { return Int(8); }
template <> inline Type type_of<int16_t>()  { return Int(16); }
template <> inline Type type_of<int32_t>()  { return Int(32); }
template <> inline Type type_of<uint8_t>()  { return UInt(8); }
template <> inline Type type_of<uint16_t>() { return UInt(16); }
template <> inline Type type_of<uint32_t>() { return UInt(32); }

// ---------------------------------------------------------------------------------------------
// type-erased dense host buffer (shared ownership, like Halide::Buffer)
// ---------------------------------------------------------------------------------------------
struct BufferData {
    Type type;
    int dims = 0;
    int extent[4] = { 0, 0, 0, 0 };
    std::vector<unsigned char> bytes;
    // a view of the trailing-index plane of a larger image (image(x, y, 2)): no bytes of its own, the
    // parent's data is read when the filter runs
    std::shared_ptr<BufferData> parent;
    size_t parent_offset = 0;          // bytes
    const unsigned char* host_data() const { return parent ? parent->host_data() + parent_offset : bytes.data(); }
    bool has_data() const { return parent ? parent->has_data() : !bytes.empty(); }
    size_t count() const
    {
        size_t n = dims ? 1 : 0;
        for (int i = 0; i < dims; ++i) n *= (size_t)extent[i];
        return n;
    }
};

class Buffer {
public:
    Buffer() {}
    Buffer(Type t, const std::vector<int>& ext) : d(std::make_shared<BufferData>())
    {
        d->type = t;
        d->dims = (int)ext.size();
        for (int i = 0; i < d->dims && i < 4; ++i) d->extent[i] = ext[i];
        d->bytes.assign(d->count() * (size_t)t.bytes(), 0);   // zero filled (halide/src/Buffer.cpp:71)
    }
    bool defined() const { return (bool)d; }
    Type type() const { return d ? d->type : Type(); }
    int dimensions() const { return d ? d->dims : 0; }
    int extent(int i) const { return (d && i < d->dims) ? d->extent[i] : 0; }
    int min(int) const { return 0; }
    void* host_ptr() const { return d ? (void*)d->bytes.data() : nullptr; }
    size_t size_in_bytes() const { return d ? d->bytes.size() : 0; }
    std::shared_ptr<BufferData> data() const { return d; }
    void copy_to_host() {}
    void copy_to_dev() {}
private:
    std::shared_ptr<BufferData> d;
};

class Realization {
public:
    Realization() {}
    explicit Realization(const std::vector<Buffer>& b) : bufs(b) {}
    size_t size() const { return bufs.size(); }
    Buffer& operator[](size_t i) { return bufs[i]; }
    const Buffer& operator[](size_t i) const { return bufs[i]; }
    operator Buffer() const { return bufs.empty() ? Buffer() : bufs[0]; }
private:
    std::vector<Buffer> bufs;
};

// ---------------------------------------------------------------------------------------------
// expressions
// ---------------------------------------------------------------------------------------------
struct ExprNode;
struct FuncDef;
class Expr {
public:
    Expr() {}
    Expr(int v);
    Expr(float v);
    Expr(double v);
    explicit Expr(std::shared_ptr<ExprNode> n) : node(n) {}
    bool defined() const { return (bool)node; }
    Type type() const;
    std::shared_ptr<ExprNode> node;
};

struct ExprNode {
    enum Kind { Const, Variable, Add, Sub, Mul, Div, Min, Max, Load, Call, Cast };
    Kind kind = Const;
    Type type = Int(32);
    double value = 0.0;                          // Const
    std::string name;                            // Variable
    std::vector<Expr> args;                      // operands / indices
    std::shared_ptr<BufferData> buffer;          // Load: the image
    std::shared_ptr<RecFilterContents> filter;   // Call: another RecFilter's result
    std::shared_ptr<FuncDef> func;               // Call: a pure Func defined by an expression (inlined by the engine)
    int tuple_index = 0;
};

// pure definition of a Func:  f(u, v) = body
struct FuncDef {
    std::string name;
    std::vector<std::string> args;
    Expr body;
};

inline Type Expr::type() const { return node ? node->type : Type(); }
inline Expr make_const(double v, Type t)
{
    auto n = std::make_shared<ExprNode>();
    n->kind = ExprNode::Const; n->value = v; n->type = t;
    return Expr(n);
}
inline Expr::Expr(int v)    { *this = make_const((double)v, Int(32)); }
inline Expr::Expr(float v)  { *this = make_const((double)v, Float(32)); }
inline Expr::Expr(double v) { *this = make_const(v, Float(64)); }

inline Type promote(const Expr& a, const Expr& b)
{
    const Type ta = a.type(), tb = b.type();
    if (ta.is_float() || tb.is_float()) return Float(std::max(ta.is_float() ? ta.bits : 32, tb.is_float() ? tb.bits : 32));
    return ta.bits >= tb.bits ? ta : tb;
}
inline Expr make_binary(ExprNode::Kind k, Expr a, Expr b)
{
    auto n = std::make_shared<ExprNode>();
    n->kind = k; n->type = promote(a, b); n->args = { a, b };
    return Expr(n);
}
inline Expr operator+(Expr a, Expr b) { return make_binary(ExprNode::Add, a, b); }
inline Expr operator-(Expr a, Expr b) { return make_binary(ExprNode::Sub, a, b); }
inline Expr operator*(Expr a, Expr b) { return make_binary(ExprNode::Mul, a, b); }
inline Expr operator/(Expr a, Expr b) { return make_binary(ExprNode::Div, a, b); }
inline Expr operator-(Expr a)         { return make_binary(ExprNode::Sub, Expr(0), a); }
inline Expr min(Expr a, Expr b)       { return make_binary(ExprNode::Min, a, b); }
inline Expr max(Expr a, Expr b)       { return make_binary(ExprNode::Max, a, b); }
inline Expr clamp(Expr a, Expr lo, Expr hi) { return max(min(a, hi), lo); }
inline Expr cast(Type t, Expr a)
{
    auto n = std::make_shared<ExprNode>();
    n->kind = ExprNode::Cast; n->type = t; n->args = { a };
    return Expr(n);
}
template <typename T> inline Expr cast(Expr a) { return cast(type_of<T>(), a); }

class Var {
public:
    Var() : n(unique_name()) {}                       // Halide gives anonymous Vars distinct names
    Var(const std::string& name) : n(name) {}
    const std::string& name() const { return n; }
    static std::string unique_name() { static int counter = 0; return "_v" + std::to_string(counter++); }
    static Var gpu_blocks() { return Var("__block_id_x"); }       // schedule granularities: accepted, ignored
    static Var gpu_threads() { return Var("__thread_id_x"); }
    operator Expr() const
    {
        auto e = std::make_shared<ExprNode>();
        e->kind = ExprNode::Variable; e->type = Int(32); e->name = n;
        return Expr(e);
    }
private:
    std::string n;
};
typedef Var RVar;
class VarOrRVar { public: VarOrRVar(const Var& v) : var(v) {} Var var; };

namespace Internal {
struct Variable {
    static Expr make(Type t, const std::string& name)
    {
        auto e = std::make_shared<ExprNode>();
        e->kind = ExprNode::Variable; e->type = t; e->name = name;
        return Expr(e);
    }
};
class Function {};
// Internal::Cast::make(type, expr), as the reference's programs spell a cast (apps/DoG/diff_gauss.cpp:73, lib/iir_coeff.cpp:180)
struct Cast {
    static Expr make(Type t, Expr e) { return Halide::cast(t, e); }
};
} // namespace Internal

class Tuple {
public:
    Tuple(Expr a) : e{ a } {}
    Tuple(Expr a, Expr b) : e{ a, b } {}
    Tuple(Expr a, Expr b, Expr c) : e{ a, b, c } {}
    explicit Tuple(const std::vector<Expr>& v) : e(v) {}
    size_t size() const { return e.size(); }
    Expr operator[](size_t i) const { return e[i]; }
    const std::vector<Expr>& as_vector() const { return e; }
private:
    std::vector<Expr> e;
};

inline Expr make_load(std::shared_ptr<BufferData> buf, const std::vector<Expr>& idx)
{
    auto n = std::make_shared<ExprNode>();
    n->kind = ExprNode::Load; n->type = buf ? buf->type : Type(); n->buffer = buf; n->args = idx;
    return Expr(n);
}

// ---------------------------------------------------------------------------------------------
// Image<T>
// ---------------------------------------------------------------------------------------------
template <typename T>
class Image {
public:
    Image() {}
    Image(int w) : buf(type_of<T>(), { w }) {}
    Image(int w, int h) : buf(type_of<T>(), { w, h }) {}
    Image(int w, int h, int c) : buf(type_of<T>(), { w, h, c }) {}
    Image(int w, int h, int c, int d) : buf(type_of<T>(), { w, h, c, d }) {}
    Image(const Buffer& b) : buf(b) { check(); }
    Image(const Realization& r) : buf(r.size() ? r[0] : Buffer()) { check(); }

    bool defined() const { return buf.defined(); }
    int dimensions() const { return buf.dimensions(); }
    int extent(int i) const { return buf.extent(i); }
    int min(int) const { return 0; }
    int width() const { return dimensions() > 0 ? extent(0) : 1; }
    int height() const { return dimensions() > 1 ? extent(1) : 1; }
    int channels() const { return dimensions() > 2 ? extent(2) : 1; }
    T* data() const { return (T*)buf.host_ptr(); }
    operator Buffer() const { return buf; }
    Buffer buffer() const { return buf; }

    T& operator()(int x, int y = 0, int z = 0, int w = 0) const
    {
        const size_t sx = (size_t)std::max(extent(0), 1), sy = (size_t)std::max(extent(1), 1), sz = (size_t)std::max(extent(2), 1);
        return data()[(size_t)x + sx * ((size_t)y + sy * ((size_t)z + sz * (size_t)w))];
    }
    Expr operator()(Expr x) const { return make_load(buf.data(), { x }); }
    Expr operator()(Expr x, Expr y) const { return make_load(buf.data(), { x, y }); }
    Expr operator()(Expr x, Expr y, Expr z) const { return make_load(buf.data(), { x, y, z }); }
    Expr operator()(Expr x, Expr y, Expr z, Expr w) const { return make_load(buf.data(), { x, y, z, w }); }
    Expr operator()(const std::vector<Expr>& idx) const { return make_load(buf.data(), idx); }
private:
    void check() const
    {
        if (buf.defined() && buf.type() != type_of<T>()) {
            std::cerr << "Image<T> constructed from a buffer of a different element type" << std::endl;
            assert(false);
        }
    }
    Buffer buf;
};

class ImageParam {
public:
    ImageParam() {}
    ImageParam(Type t, int dims, const std::string& name = "") : ty(t), nd(dims), nm(name) {}
    template <typename T> void set(const Image<T>& im) { buf = im.buffer(); }
    void set(const Buffer& b) { buf = b; }
    Buffer get() const { return buf; }
    Type type() const { return ty; }
    int dimensions() const { return nd; }
    Expr operator()(Expr x) const { return make_load(buf.data(), { x }); }
    Expr operator()(Expr x, Expr y) const { return make_load(buf.data(), { x, y }); }
    Expr operator()(Expr x, Expr y, Expr z) const { return make_load(buf.data(), { x, y, z }); }
    Expr operator()(Expr x, Expr y, Expr z, Expr w) const { return make_load(buf.data(), { x, y, z, w }); }
private:
    Type ty; int nd = 0; std::string nm; Buffer buf;
};

// ---------------------------------------------------------------------------------------------
// Func: a handle on a RecFilter's result (what RecFilter::as_func() returns).  Calling it builds a
// Call expression; the schedule methods exist so that user code compiles and are ignored -- the
// launch planner decides how kernels run.
// ---------------------------------------------------------------------------------------------
class FuncRefExpr {
public:
    FuncRefExpr(std::shared_ptr<RecFilterContents> f, const std::vector<Expr>& a) : filter(f), args(a) {}
    FuncRefExpr(std::shared_ptr<RecFilterContents> f, std::shared_ptr<FuncDef> d, const std::vector<Expr>& a)
        : filter(f), def(d), args(a) {}
    operator Expr() const;
    Expr operator[](int i) const;
    // pure definition  f(u, v) = expr  (the arguments must be plain Vars)
    void operator=(Expr body)
    {
        if (filter || !def) { std::cerr << "Func: only a Func that is not a RecFilter result can be defined" << std::endl; assert(false); }
        def->args.clear();
        for (const Expr& a : args) {
            if (!a.defined() || a.node->kind != ExprNode::Variable) {
                std::cerr << "Func: the arguments of a pure definition must be Vars" << std::endl; assert(false);
            }
            def->args.push_back(a.node->name);
        }
        def->body = body;
    }
    void operator=(const FuncRefExpr& other) { *this = Expr(other); }
    std::shared_ptr<RecFilterContents> filter;
    std::shared_ptr<FuncDef> def;
    std::vector<Expr> args;
};
typedef FuncRefExpr FuncRefVar;

class Func {
public:
    Func() : def(std::make_shared<FuncDef>()) {}
    explicit Func(const std::string& n) : def(std::make_shared<FuncDef>()), nm(n) { def->name = n; }
    explicit Func(std::shared_ptr<RecFilterContents> f, const std::string& n = "") : filter(f), nm(n) {}
    const std::string& name() const { return nm; }
    bool defined() const { return (bool)filter || (def && def->body.defined()); }
    FuncRefExpr operator()(Expr x) const { return FuncRefExpr(filter, def, { x }); }
    FuncRefExpr operator()(Expr x, Expr y) const { return FuncRefExpr(filter, def, { x, y }); }
    FuncRefExpr operator()(Expr x, Expr y, Expr z) const { return FuncRefExpr(filter, def, { x, y, z }); }
    FuncRefExpr operator()(const std::vector<Expr>& a) const { return FuncRefExpr(filter, def, a); }
    // accepted and ignored schedule directives
    Func& compute_root() { return *this; }
    Func& compute_at(Func, Var) { return *this; }
    Func& split(Var, Var, Var, int) { return *this; }
    Func& unroll(Var) { return *this; }
    Func& vectorize(Var, int = 0) { return *this; }
    Func& parallel(Var) { return *this; }
    Func& reorder(Var, Var) { return *this; }
    Func& reorder(Var, Var, Var) { return *this; }
    Func& reorder(Var, Var, Var, Var) { return *this; }
    Func& reorder(Var, Var, Var, Var, Var) { return *this; }
    Func& gpu(Var, Var, Var, Var) { return *this; }
    Func& gpu_tile(Var, Var, int, int) { return *this; }
    Func& bound(Var, Expr, Expr) { return *this; }
    std::shared_ptr<RecFilterContents> filter;
    std::shared_ptr<FuncDef> def;                // pure definition (when the Func is not a RecFilter result)
private:
    std::string nm;
};

inline FuncRefExpr::operator Expr() const
{
    auto n = std::make_shared<ExprNode>();
    n->kind = ExprNode::Call; n->filter = filter; n->args = args;
    if (!filter) n->func = def;
    n->type = Float(32);          // refined by RecFilter::define from the callee's type
    return Expr(n);
}
inline Expr FuncRefExpr::operator[](int i) const
{
    Expr e = *this;
    e.node->tuple_index = i;
    return e;
}

struct Target {
    enum Feature { CUDA, CUDACapability100 };
    bool has_gpu_feature() const { return true; }      // always: there is no CPU execution path
    bool has_feature(Feature) const { return true; }
    std::string to_string() const { return "b200-sm_100a"; }
};
inline Target get_jit_target_from_environment() { return Target(); }

} // namespace Halide

#endif // RECFILTER_B200_HALIDE_SHIM_H_
