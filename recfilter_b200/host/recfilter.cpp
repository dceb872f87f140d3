/*
 * recfilter.cpp -- the RecFilter operator surface (include/recfilter.h) on top of the C ABI of
 * the B200 engine (include/recfilter_b200.h).  Host C++ only: no CUDA headers, no Halide.
 *
 * Mirrors /root/reference/lib/recfilter.cpp (define :192-248, add_filter :260-392,
 * set_clamped_image_border :252-258, realize :984-989, profile :991-1016, as_func :886-903),
 * lib/split.cpp:1850-2111 (split: here it only records the tile hint -- tiling is the
 * planner's business), lib/reorder.cpp (cascade :28-229, overlap_to_higher_order_filter
 * :231-381) and lib/recfilter_utils.cpp (Arguments :31-112, operator+/- :274-275).
 *
 * The filter is a value: dimensions, element type, the array it filters (an image, or the
 * result of another RecFilter), the scan list in add_filter order, the border mode.  realize()
 * resolves the input chain, creates (once) a launch plan per filter through rf_plan_create and
 * runs the plans on device-resident buffers.  There is no CPU execution path.
 */
#include <recfilter.h>
#include <iir_coeff.h>
#include <recfilter_b200.h>

#include <cassert>
#include <climits>
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>

using namespace Halide;
using std::cerr;
using std::endl;
using std::string;
using std::vector;

// ---------------------------------------------------------------------------------------------
// filter state
// ---------------------------------------------------------------------------------------------
struct ScanDef {
    int dim;
    bool causal;
    vector<float> coeff;            // {b0, a1..ar}
};

struct RecFilterContents {
    string name;
    vector<RecFilterDim> dims;
    Type type;
    bool defined = false;
    bool clamped = false;
    bool tiled = false;
    vector<Expr> rhs;               // pure definition as written by the user
    // resolved input: an image (dense host buffer) or the result of another filter
    std::shared_ptr<BufferData> src_image;
    std::shared_ptr<RecFilterContents> src_filter;
    // pointwise linear stencil of the input applied before the scans (empty: the input itself):
    // box-filter finite differencing and similar epilogues (apps/box/box_filter.h:36-39)
    vector<rf_tap> stencil;
    float stencil_scale = 1.0f;     // common factor of the taps, applied after the sum
    // second source of a stencil: an image read beside the result of src_filter
    // (unsharp mask: (1+w)*image - w*blur, apps/usm/unsharp_mask_naive.cpp:61); taps with source == 1 read it
    std::shared_ptr<BufferData> side_image;
    // ... or the result of a second filter (difference of Gaussians: two channels of a Tuple filter subtracted,
    // apps/DoG/diff_gauss.cpp:96): taps with source == 1 read it.  At most one of side_image / side_filter is set.
    std::shared_ptr<RecFilterContents> side_filter;
    void* dev_side = nullptr;
    vector<ScanDef> scans;
    std::map<string, int> tiles;    // split() hints
    rf_plan* plan = nullptr;
    string plan_sig;                // what the plan was built for: scans (its own, or a fused cascade's), extents, border, tiles
    rf_mgpu* mgpu = nullptr;        // RECFILTER_GPUS=n: the same filter cut over n GPUs (rf_mgpu_*)
    string mgpu_sig;
    bool mgpu_refused = false;      // the engine declined to shard this filter: said once, then one GPU
    bool epi_refused = false;       // the engine declined to fuse this pointwise definition into its filter's last store
    // Tuple (multi-output) filters, lib/recfilter.cpp:68-74,197-203: every Tuple element is filtered
    // independently with the same scans -- one channel filter per element, kept in step by sync_channels()
    vector<std::shared_ptr<RecFilterContents>> channels;
    void* dev_out = nullptr;        // device result buffer, reused across realize()/profile()
    void* dev_in = nullptr;         // the input image on the device while profile() keeps inputs resident
    void* dev_tmp = nullptr;        // stencil output when scans follow
    ~RecFilterContents()
    {
        if (plan) rf_plan_destroy(plan);
        if (mgpu) rf_mgpu_destroy(mgpu);
        if (dev_out) rf_free(dev_out);
        if (dev_in) rf_free(dev_in);
        if (dev_tmp) rf_free(dev_tmp);
        if (dev_side) rf_free(dev_side);
    }
    size_t count() const { size_t n = 1; for (const auto& d : dims) n *= (size_t)d.num_pixels(); return n; }
    int dim_index(const string& var) const
    {
        for (size_t i = 0; i < dims.size(); ++i) if (dims[i].var().name() == var) return (int)i;
        return -1;
    }
};

namespace {

[[noreturn]] void die(const string& msg)
{
    cerr << msg << endl;
    assert(false);
    abort();                        // asserts are live in the reference build; keep failing without them
}

void engine_check(int rc, const char* what)
{
    if (rc != RF_OK) die(string(what) + " failed: " + rf_last_error());
}

int engine_dtype(Type t)
{
    if (t == Float(32)) return RF_F32;
    if (t == Int(32)) return RF_I32;
    if (t == UInt(32)) return RF_U32;
    if (t == Int(16)) return RF_I16;
    if (t == UInt(16)) return RF_U16;
    if (t == Int(8)) return RF_I8;
    if (t == UInt(8)) return RF_U8;
    die("RecFilter: element type not supported by the B200 engine (float32 and 8/16/32-bit integers are)");
}

// is `e` the variable `name`, possibly wrapped in clamp(name, 0, extent-1) (an identity inside the domain)?
bool is_identity_index(const Expr& e, const string& name, int extent)
{
    if (!e.defined()) return false;
    const ExprNode& n = *e.node;
    if (n.kind == ExprNode::Variable) return n.name == name;
    if (n.kind == ExprNode::Max && n.args.size() == 2) {          // clamp(x, lo, hi) == max(min(x, hi), lo)
        const Expr &inner = n.args[0], &lo = n.args[1];
        if (lo.node->kind != ExprNode::Const || lo.node->value != 0.0) return false;
        if (inner.node->kind != ExprNode::Min || inner.node->args.size() != 2) return false;
        const Expr &v = inner.node->args[0], &hi = inner.node->args[1];
        if (hi.node->kind != ExprNode::Const || hi.node->value < double(extent - 1)) return false;
        return is_identity_index(v, name, extent);
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// definitions that are a small linear stencil of ONE image / filter result:
//   D(x,y) = (f(clamp(x+B,..), clamp(y+B,..)) - f(..) + ..) / area      (apps/box/box_filter.h:36-39)
// possibly through pure helper Funcs (fA(u,v) = ..., box_filter.h:122-131), which are inlined.
// ---------------------------------------------------------------------------------------------
void sync_channels(RecFilterContents& c);

struct IndexMap {                  // clamp(var_dim + off, lo, hi)
    int dim = -1;
    long long off = 0, lo = INT32_MIN, hi = INT32_MAX;
};
struct LinTap {
    double w = 0.0;
    std::shared_ptr<BufferData> image;
    std::shared_ptr<RecFilterContents> filter;
    vector<IndexMap> idx;
};
typedef std::map<string, IndexMap> IndexEnv;

bool const_value(const Expr& e, double& v)
{
    if (!e.defined()) return false;
    const ExprNode& n = *e.node;
    double a, b;
    switch (n.kind) {
    case ExprNode::Const: v = n.value; return true;
    case ExprNode::Cast:  return const_value(n.args[0], v);
    case ExprNode::Add: if (const_value(n.args[0], a) && const_value(n.args[1], b)) { v = a + b; return true; } return false;
    case ExprNode::Sub: if (const_value(n.args[0], a) && const_value(n.args[1], b)) { v = a - b; return true; } return false;
    case ExprNode::Mul: if (const_value(n.args[0], a) && const_value(n.args[1], b)) { v = a * b; return true; } return false;
    case ExprNode::Div: if (const_value(n.args[0], a) && const_value(n.args[1], b) && b != 0.0) { v = a / b; return true; } return false;
    default: return false;
    }
}

IndexMap eval_index(const RecFilterContents& c, const Expr& e, const IndexEnv& env)
{
    if (!e.defined()) die("RecFilter " + c.name + ": undefined index expression");
    const ExprNode& n = *e.node;
    double k;
    auto shift = [](IndexMap m, long long d) {
        m.off += d;
        if (m.lo != INT32_MIN) m.lo += d;
        if (m.hi != INT32_MAX) m.hi += d;
        return m;
    };
    switch (n.kind) {
    case ExprNode::Variable: {
        auto it = env.find(n.name);
        if (it != env.end()) return it->second;
        IndexMap m; m.dim = c.dim_index(n.name);
        if (m.dim < 0) die("RecFilter " + c.name + ": variable " + n.name + " in an index is not a dimension of the filter");
        return m;
    }
    case ExprNode::Cast: return eval_index(c, n.args[0], env);
    case ExprNode::Add:
        if (const_value(n.args[1], k)) return shift(eval_index(c, n.args[0], env), (long long)k);
        if (const_value(n.args[0], k)) return shift(eval_index(c, n.args[1], env), (long long)k);
        break;
    case ExprNode::Sub:
        if (const_value(n.args[1], k)) return shift(eval_index(c, n.args[0], env), -(long long)k);
        break;
    case ExprNode::Min:
    case ExprNode::Max: {
        const bool c1 = const_value(n.args[1], k);
        if (!c1 && !const_value(n.args[0], k)) break;
        IndexMap m = eval_index(c, n.args[c1 ? 0 : 1], env);
        const long long kk = (long long)k;
        if (n.kind == ExprNode::Min) { m.lo = m.lo == INT32_MIN ? m.lo : std::min(m.lo, kk); m.hi = std::min(m.hi, kk); }
        else                         { m.lo = std::max(m.lo, kk); m.hi = m.hi == INT32_MAX ? m.hi : std::max(m.hi, kk); }
        return m;
    }
    default: break;
    }
    die("RecFilter " + c.name + ": index expressions must be a dimension plus a constant, optionally clamped "
        "(min / max / clamp with constants)");
}

// e -> list of taps (+ a constant term that must vanish)
void linearize(const RecFilterContents& c, const Expr& e, const IndexEnv& env, double scale, vector<LinTap>& taps,
               double& constant, int depth = 0)
{
    if (!e.defined()) die("RecFilter " + c.name + ": undefined expression in the definition");
    if (depth > 32) die("RecFilter " + c.name + ": Func definitions nest too deeply (recursive definition?)");
    const ExprNode& n = *e.node;
    double k;
    if (const_value(e, k)) { constant += scale * k; return; }
    switch (n.kind) {
    case ExprNode::Cast: linearize(c, n.args[0], env, scale, taps, constant, depth); return;
    case ExprNode::Add:
        linearize(c, n.args[0], env, scale, taps, constant, depth);
        linearize(c, n.args[1], env, scale, taps, constant, depth);
        return;
    case ExprNode::Sub:
        linearize(c, n.args[0], env, scale, taps, constant, depth);
        linearize(c, n.args[1], env, -scale, taps, constant, depth);
        return;
    case ExprNode::Mul:
        if (const_value(n.args[0], k)) { linearize(c, n.args[1], env, scale * k, taps, constant, depth); return; }
        if (const_value(n.args[1], k)) { linearize(c, n.args[0], env, scale * k, taps, constant, depth); return; }
        break;
    case ExprNode::Div:
        if (const_value(n.args[1], k) && k != 0.0) { linearize(c, n.args[0], env, scale / k, taps, constant, depth); return; }
        break;
    case ExprNode::Load:
    case ExprNode::Call: {
        if (n.kind == ExprNode::Call && !n.filter) {              // pure Func: inline its definition
            if (!n.func || !n.func->body.defined()) die("RecFilter " + c.name + ": call of an undefined Func");
            if (n.func->args.size() != n.args.size()) die("RecFilter " + c.name + ": Func called with the wrong number of arguments");
            IndexEnv inner;
            for (size_t i = 0; i < n.args.size(); ++i) inner[n.func->args[i]] = eval_index(c, n.args[i], env);
            linearize(c, n.func->body, inner, scale, taps, constant, depth + 1);
            return;
        }
        LinTap t; t.w = scale; t.image = n.kind == ExprNode::Load ? n.buffer : nullptr; t.filter = n.filter;
        if (t.filter && !t.filter->channels.empty()) {              // F(x,y)[i] of a Tuple filter
            if (n.tuple_index < 0 || n.tuple_index >= (int)t.filter->channels.size())
                die("RecFilter " + c.name + ": Tuple index out of range in the definition");
            sync_channels(*t.filter);
            t.filter = t.filter->channels[n.tuple_index];
        }
        size_t nidx = n.args.size();
        if (t.image && nidx > c.dims.size() && (int)nidx == t.image->dims) {
            // image(x, y, 2): trailing constant indices select one plane of a larger image
            size_t off = 0, stride = 1;
            for (size_t i = 0; i < nidx; ++i) {
                if (i >= c.dims.size()) {
                    double k;
                    if (!const_value(n.args[i], k) || k < 0 || k >= t.image->extent[i])
                        die("RecFilter " + c.name + ": indices beyond the filter's dimensions must be constants inside the image");
                    off += (size_t)k * stride;
                }
                stride *= (size_t)t.image->extent[i];
            }
            // one view per plane, so that two taps of the same plane are "the same image"; the table holds weak
            // references (a view lives as long as a filter reads it, and keeps its parent image alive that long)
            static std::map<std::pair<BufferData*, size_t>, std::weak_ptr<BufferData>> views;
            for (auto it = views.begin(); it != views.end(); ) it = it->second.expired() ? views.erase(it) : std::next(it);
            auto key = std::make_pair(t.image.get(), off);
            std::shared_ptr<BufferData> v = views.count(key) ? views[key].lock() : nullptr;
            if (!v) {
                v = std::make_shared<BufferData>();
                v->type = t.image->type; v->dims = (int)c.dims.size();
                for (size_t i = 0; i < c.dims.size(); ++i) v->extent[i] = t.image->extent[i];
                v->parent = t.image; v->parent_offset = off * (size_t)t.image->type.bytes();
                views[key] = v;
            }
            t.image = v;
            nidx = c.dims.size();
        }
        for (size_t i = 0; i < nidx; ++i) t.idx.push_back(eval_index(c, n.args[i], env));
        taps.push_back(t);
        return;
    }
    default: break;
    }
    die("RecFilter " + c.name + ": the definition must be linear in the image / filter it reads (sums, differences and "
        "constant factors of taps); general expressions are not supported by the B200 engine");
}

// ---------------------------------------------------------------------------------------------
// cascade fusion.  A filter whose input is another filter's result, unchanged, continues that filter's
// scan list: running "x scans -> store -> load -> y scans" (cascade_by_dimension, apps/gaussian/
// gaussian_filter_3x_3y.cpp:45-52) is the same computation as one filter with all the scans, and the
// planner can then fuse them into one pass instead of one HBM round-trip per link (the reference pays a
// full round-trip per link, lib/reorder.cpp:28-176).  unit_first(c) is the first filter of the maximal
// fusable run that ends in c.  RECFILTER_NO_CHAIN_FUSION=1 turns the fusion off.
// ---------------------------------------------------------------------------------------------
bool fusable_with_parent(const RecFilterContents& f)
{
    static const bool off = getenv("RECFILTER_NO_CHAIN_FUSION") && atoi(getenv("RECFILTER_NO_CHAIN_FUSION")) != 0;
    if (off || !f.src_filter || !f.stencil.empty()) return false;
    const RecFilterContents& p = *f.src_filter;
    if (!p.defined || p.clamped != f.clamped || !(p.type == f.type) || p.dims.size() != f.dims.size()) return false;
    for (size_t i = 0; i < f.dims.size(); ++i)
        if (p.dims[i].num_pixels() != f.dims[i].num_pixels()) return false;
    return true;
}
RecFilterContents& unit_first(RecFilterContents& c)
{
    // the fused tile kernels take up to 4 scans per dimension: a longer run would fall back to the generic engine,
    // which is slower than one more fused pass
    const int cap = 4;
    int per_dim[RF_MAX_DIMS] = { 0, 0, 0, 0 };
    auto add = [&](const RecFilterContents& g) {
        int tmp[RF_MAX_DIMS];
        for (int d = 0; d < RF_MAX_DIMS; ++d) tmp[d] = per_dim[d];
        for (const ScanDef& sc : g.scans) if (sc.dim >= 0 && sc.dim < RF_MAX_DIMS && ++tmp[sc.dim] > cap) return false;
        for (int d = 0; d < RF_MAX_DIMS; ++d) per_dim[d] = tmp[d];
        return true;
    };
    RecFilterContents* f = &c;
    add(c);
    while (fusable_with_parent(*f) && add(*f->src_filter)) f = f->src_filter.get();
    return *f;
}
vector<ScanDef> unit_scans(RecFilterContents& first, RecFilterContents& last)
{
    vector<RecFilterContents*> run;
    for (RecFilterContents* f = &last; ; f = f->src_filter.get()) { run.push_back(f); if (f == &first) break; }
    vector<ScanDef> scans;
    for (auto it = run.rbegin(); it != run.rend(); ++it) scans.insert(scans.end(), (*it)->scans.begin(), (*it)->scans.end());
    return scans;
}

// descriptor of the filter c with the scan list `scans`, and a signature of everything a plan depends on
string fill_desc(RecFilterContents& c, const vector<ScanDef>& scans, rf_desc& d)
{
    std::memset(&d, 0, sizeof(d));
    if (c.dims.size() > RF_MAX_DIMS) die("RecFilter: at most 4 dimensions are supported");
    if (scans.size() > RF_MAX_SCANS) die("RecFilter: too many scans");
    d.ndim = (int)c.dims.size();
    for (int i = 0; i < d.ndim; ++i) d.extent[i] = c.dims[i].num_pixels();
    d.dtype = engine_dtype(c.type);
    d.border = c.clamped ? RF_BORDER_CLAMP : RF_BORDER_ZERO;
    d.nscans = (int)scans.size();
    for (int s = 0; s < d.nscans; ++s) {
        const ScanDef& sc = scans[s];
        if ((int)sc.coeff.size() - 1 > RF_MAX_ORDER) die("RecFilter: filter order above 32 is not supported");
        d.scans[s].dim = sc.dim;
        d.scans[s].causal = sc.causal ? 1 : 0;
        d.scans[s].order = (int)sc.coeff.size() - 1;
        for (size_t k = 0; k < sc.coeff.size(); ++k) d.scans[s].coeff[k] = sc.coeff[k];
    }
    for (int i = 0; i < d.ndim; ++i) {
        auto it = c.tiles.find(c.dims[i].var().name());
        d.opt.tile[i] = it == c.tiles.end() ? 0 : it->second;      // a hint: the planner picks its own register tile
    }
    d.opt.fuse_dims = -1;
    d.opt.shard_dim = -1;
    std::ostringstream sig;
    sig << d.dtype << (c.clamped ? 'c' : 'z');
    for (int i = 0; i < d.ndim; ++i) sig << d.extent[i] << '/' << d.opt.tile[i] << 'x';
    for (const ScanDef& sc : scans) { sig << sc.dim << (sc.causal ? '+' : '-'); for (float v : sc.coeff) sig << v << ','; sig << ';'; }
    return sig.str();
}

void build_plan(RecFilterContents& c, const vector<ScanDef>& scans)
{
    rf_desc d;
    const string sig = fill_desc(c, scans, d);
    if (c.plan && c.plan_sig == sig) return;
    if (c.plan) { rf_plan_destroy(c.plan); c.plan = nullptr; }
    engine_check(rf_plan_create(&d, &c.plan), "rf_plan_create");
    c.plan_sig = sig;
}
void build_plan(RecFilterContents& c) { build_plan(c, unit_scans(unit_first(c), c)); }

// Device buffer holding the input of `c` (uploads an image, or evaluates the upstream filter).
// `owned` tells the caller whether it must free the buffer.
void* evaluate_device(RecFilterContents& c);

// dense device copy of the [0, extent) box of a host image (which may be larger than the filter domain)
// one element of an image of type `t`, as a double
static double load_as_double(const unsigned char* p, const Type& t)
{
    if (t.is_float()) return t.bits == 32 ? (double)*reinterpret_cast<const float*>(p) : *reinterpret_cast<const double*>(p);
    if (t.is_int()) {
        switch (t.bits) { case 8: return *reinterpret_cast<const int8_t*>(p); case 16: return *reinterpret_cast<const int16_t*>(p);
                          default: return *reinterpret_cast<const int32_t*>(p); }
    }
    switch (t.bits) { case 8: return *reinterpret_cast<const uint8_t*>(p); case 16: return *reinterpret_cast<const uint16_t*>(p);
                      default: return *reinterpret_cast<const uint32_t*>(p); }
}

void* upload_image(RecFilterContents& c, const BufferData& b)
{
    const size_t bytes = c.count() * (size_t)c.type.bytes();
    void* dev = nullptr;
    engine_check(rf_malloc(&dev, bytes), "rf_malloc");
    bool same = b.dims == (int)c.dims.size();
    for (int i = 0; same && i < b.dims; ++i) same = b.extent[i] == c.dims[i].num_pixels();
    const bool convert = !(b.type == c.type);             // cast<float>(image16(x, y)): the filter type is the RHS type
    if (same && !convert) {
        engine_check(rf_memcpy_h2d(dev, b.host_data(), bytes), "rf_memcpy_h2d");
    } else {
        // the image is larger than the filter domain and / or of another type: pack (and convert) the [0, extent) box
        const int eb = c.type.bytes(), sb = b.type.bytes();
        vector<unsigned char> packed(bytes);
        int ext[4] = { 1, 1, 1, 1 }, bext[4] = { 1, 1, 1, 1 };
        for (size_t i = 0; i < c.dims.size(); ++i) { ext[i] = c.dims[i].num_pixels(); bext[i] = b.extent[i]; }
        size_t o = 0;
        for (int w = 0; w < ext[3]; ++w)
            for (int z = 0; z < ext[2]; ++z)
                for (int y = 0; y < ext[1]; ++y) {
                    const size_t src = (((size_t)w * bext[2] + z) * bext[1] + y) * (size_t)bext[0];
                    if (!convert) {
                        std::memcpy(&packed[o * eb], b.host_data() + src * eb, (size_t)ext[0] * eb);
                    } else {                                // integer image -> float filter (the only conversion define() accepts)
                        float* dstp = reinterpret_cast<float*>(&packed[o * eb]);
                        for (int x = 0; x < ext[0]; ++x) dstp[x] = (float)load_as_double(b.host_data() + (src + x) * sb, b.type);
                    }
                    o += (size_t)ext[0];
                }
        engine_check(rf_memcpy_h2d(dev, packed.data(), bytes), "rf_memcpy_h2d");
    }
    return dev;
}

// profile(): images stay resident over the timed iterations (kernels only are timed, lib/recfilter.cpp:995-1011), also
// the root image of a second-source filter chain that run_unit evaluates on the way (apps/DoG)
bool g_resident_inputs = false;
vector<RecFilterContents*> g_resident_owners;

void* input_device(RecFilterContents& c, bool& owned)
{
    if (c.src_filter) { owned = false; return evaluate_device(*c.src_filter); }
    if (!c.src_image) die("RecFilter " + c.name + ": the input image is not set (ImageParam::set was not called?)");
    if (g_resident_inputs) {
        if (!c.dev_in) { c.dev_in = upload_image(c, *c.src_image); g_resident_owners.push_back(&c); }
        owned = false;
        return c.dev_in;
    }
    owned = true;
    return upload_image(c, *c.src_image);
}
void release_resident_inputs()
{
    g_resident_inputs = false;
    if (!g_resident_owners.empty()) engine_check(rf_synchronize(), "rf_synchronize");
    for (RecFilterContents* o : g_resident_owners)
        if (o->dev_in) { engine_check(rf_free(o->dev_in), "rf_free"); o->dev_in = nullptr; }
    g_resident_owners.clear();
}

// run the unit first..c on the device buffer `in` (the input of `first`): the stencil of first's definition
// (if any), then every scan of the run in one plan
void* run_unit(RecFilterContents& first, RecFilterContents& c, const void* in, bool refresh_side = true)
{
    const size_t bytes = c.count() * (size_t)c.type.bytes();
    const vector<ScanDef> scans = unit_scans(first, c);
    if (!c.dev_out) engine_check(rf_malloc(&c.dev_out, bytes), "rf_malloc");
    if (!first.stencil.empty()) {
        RecFilterContents& f = first;
        int64_t ext[RF_MAX_DIMS] = { 1, 1, 1, 1 };
        for (size_t i = 0; i < c.dims.size(); ++i) ext[i] = c.dims[i].num_pixels();
        void* dst = c.dev_out;
        if (!scans.empty()) {
            if (!c.dev_tmp) engine_check(rf_malloc(&c.dev_tmp, bytes), "rf_malloc");
            dst = c.dev_tmp;
        }
        if (f.side_image && (refresh_side || !f.dev_side)) {
            // uploaded for every evaluation, like the primary image: the caller may have changed it between realize() calls
            // (profile() keeps it resident over its timed iterations, like the primary image)
            if (f.dev_side) { engine_check(rf_synchronize(), "rf_synchronize"); engine_check(rf_free(f.dev_side), "rf_free"); f.dev_side = nullptr; }
            f.dev_side = upload_image(f, *f.side_image);
        }
        // the second source: the resident side image, or the result of the second filter (evaluated now, into that
        // filter's own device buffer)
        const void* in2 = f.side_filter ? evaluate_device(*f.side_filter) : f.dev_side;
        engine_check(rf_stencil_execute((int)c.dims.size(), ext, engine_dtype(c.type), (int)f.stencil.size(), f.stencil.data(),
                                        f.stencil_scale, in, in2, dst, nullptr), "rf_stencil_execute");
        if (scans.empty()) return c.dev_out;
        in = dst;
    }
    build_plan(c, scans);
    engine_check(rf_plan_execute(c.plan, in, c.dev_out, nullptr), "rf_plan_execute");
    return c.dev_out;
}


// ---------------------------------------------------------------------------------------------
// RECFILTER_GPUS=n: realize() / profile() cut the filter over n GPUs of this process (rf_mgpu_*: independent parts when
// the outermost dimension carries no scans, strips with one order-r tail exchange over NVLink otherwise).  It applies
// to a filter that is ONE plan on ONE input image of exactly its extents (every in-scope app: apps/gaussian,
// apps/summed_table, apps/audio ...); anything else -- and anything the engine declines, e.g. an extent that
// does not divide -- runs on one GPU as before, with a note on stderr.
// ---------------------------------------------------------------------------------------------
int requested_gpus()
{
    static const int n = getenv("RECFILTER_GPUS") ? std::max(1, atoi(getenv("RECFILTER_GPUS"))) : 1;
    return n;
}
rf_mgpu* mgpu_for(RecFilterContents& c)
{
    if (requested_gpus() < 2 || c.mgpu_refused || !c.channels.empty()) return nullptr;
    RecFilterContents& first = unit_first(c);
    if (first.src_filter || !first.stencil.empty() || !first.src_image || !first.src_image->has_data()) return nullptr;
    const BufferData& b = *first.src_image;
    if (b.dims != (int)c.dims.size()) return nullptr;
    for (int i = 0; i < b.dims; ++i) if (b.extent[i] != c.dims[i].num_pixels()) return nullptr;
    rf_desc d;
    const string sig = fill_desc(c, unit_scans(first, c), d);
    if (c.mgpu && c.mgpu_sig == sig) return c.mgpu;
    if (c.mgpu) { rf_mgpu_destroy(c.mgpu); c.mgpu = nullptr; }
    const int rc = rf_mgpu_create(&d, requested_gpus(), &c.mgpu);
    if (rc != RF_OK) {
        cerr << "RecFilter " << c.name << ": RECFILTER_GPUS=" << requested_gpus() << " not used (" << rf_mgpu_last_error()
             << "); running on one GPU" << endl;
        c.mgpu = nullptr; c.mgpu_refused = true;
        return nullptr;
    }
    c.mgpu_sig = sig;
    return c.mgpu;
}

// ---------------------------------------------------------------------------------------------
// epilogue fusion.  A definition without scans that is  a * image(x, y) + b * F(x, y)  -- F a filter (or fused run of
// filters) whose input is that same image, unchanged -- is the unsharp mask of apps/usm/unsharp_mask_optimized.cpp:
// 61-66, where the reference merges the blur's last stage into USM with RecFilter::compute_at.  Here the combination
// is applied by the store of the filter's last kernel (rf_options.epilogue): the blurred image never travels to HBM
// and back.  Anything the engine declines (RF_EUNSUPPORTED) runs as the separate stencil kernel, as before.
// RECFILTER_NO_EPILOGUE_FUSION=1 turns the fusion off.
// ---------------------------------------------------------------------------------------------
struct EpilogueMatch {
    RecFilterContents *first = nullptr, *last = nullptr;      // the filter run whose store takes the epilogue
    float a_in = 0.0f, a_out = 0.0f;
};
bool epilogue_match(RecFilterContents& c, EpilogueMatch& m)
{
    static const bool off = getenv("RECFILTER_NO_EPILOGUE_FUSION") && atoi(getenv("RECFILTER_NO_EPILOGUE_FUSION")) != 0;
    if (off || c.epi_refused || !c.scans.empty() || !c.src_filter || !c.side_image || c.stencil.size() != 2) return false;
    if (!(c.type == Float(32)) || !c.side_image->has_data()) return false;
    const rf_tap *tf = nullptr, *ti = nullptr;
    for (const rf_tap& t : c.stencil) {
        for (int d = 0; d < RF_MAX_DIMS; ++d) if (t.offset[d] != 0) return false;
        (t.source ? ti : tf) = &t;
    }
    if (!tf || !ti) return false;
    RecFilterContents& last = *c.src_filter;
    if (!last.defined || !last.channels.empty() || last.scans.empty() || !(last.type == c.type) || last.dims.size() != c.dims.size()) return false;
    RecFilterContents& first = unit_first(last);
    if (first.src_filter || !first.stencil.empty() || first.src_image != c.side_image) return false;
    const BufferData& b = *c.side_image;
    if (!(b.type == c.type) || b.dims != (int)c.dims.size()) return false;
    for (size_t i = 0; i < c.dims.size(); ++i)
        if (b.extent[i] != c.dims[i].num_pixels() || last.dims[i].num_pixels() != c.dims[i].num_pixels()) return false;
    m.first = &first; m.last = &last;
    m.a_in = c.stencil_scale * ti->weight; m.a_out = c.stencil_scale * tf->weight;
    return true;
}
// the matched run with the epilogue on `in` (the image on the device); null when the engine declines the fusion
void* run_epilogue_unit(RecFilterContents& c, const EpilogueMatch& m, const void* in)
{
    rf_desc d;
    std::ostringstream sig;
    sig << fill_desc(*m.last, unit_scans(*m.first, *m.last), d) << "|epilogue " << m.a_in << ' ' << m.a_out;
    d.opt.epilogue = 1; d.opt.epi_in = m.a_in; d.opt.epi_out = m.a_out;
    if (!(c.plan && c.plan_sig == sig.str())) {
        if (c.plan) { rf_plan_destroy(c.plan); c.plan = nullptr; }
        const int rc = rf_plan_create(&d, &c.plan);
        if (rc == RF_EUNSUPPORTED) { c.plan = nullptr; c.epi_refused = true; return nullptr; }
        engine_check(rc, "rf_plan_create");
        c.plan_sig = sig.str();
    }
    if (!c.dev_out) engine_check(rf_malloc(&c.dev_out, c.count() * (size_t)c.type.bytes()), "rf_malloc");
    engine_check(rf_plan_execute(c.plan, in, c.dev_out, nullptr), "rf_plan_execute");
    return c.dev_out;
}

void* evaluate_device(RecFilterContents& c)
{
    if (!c.defined) die("RecFilter " + c.name + " is used before it is defined");
    EpilogueMatch epi;
    if (epilogue_match(c, epi)) {
        void* img = upload_image(c, *c.side_image);
        void* out = run_epilogue_unit(c, epi, img);
        engine_check(rf_synchronize(), "rf_synchronize"); engine_check(rf_free(img), "rf_free");
        if (out) return out;
    }
    RecFilterContents& first = unit_first(c);
    bool owned = false;
    void* in = input_device(first, owned);
    void* out = run_unit(first, c, in);
    if (owned) { engine_check(rf_synchronize(), "rf_synchronize"); engine_check(rf_free(in), "rf_free"); }
    return out;
}

// a Tuple filter's channels follow its scans, tiling hints and border mode
void sync_channels(RecFilterContents& c)
{
    for (auto& ch : c.channels) {
        ch->scans = c.scans;
        ch->tiles = c.tiles;
        ch->clamped = c.clamped;
    }
}

// every filter of the chain ending in c as declared, upstream first
void collect_chain(RecFilterContents& c, vector<RecFilterContents*>& chain)
{
    if (c.src_filter) collect_chain(*c.src_filter, chain);
    chain.push_back(&c);
}
// the execution units (fused runs) of the chain ending in c, upstream first: (first, last) pairs
void collect_units(RecFilterContents& c, vector<std::pair<RecFilterContents*, RecFilterContents*>>& units)
{
    RecFilterContents& first = unit_first(c);
    if (first.src_filter) collect_units(*first.src_filter, units);
    units.push_back({ &first, &c });
}

} // namespace

// ---------------------------------------------------------------------------------------------
// operators, statics
// ---------------------------------------------------------------------------------------------
RecFilterDimAndCausality operator+(RecFilterDim x) { return RecFilterDimAndCausality(x, true); }
RecFilterDimAndCausality operator-(RecFilterDim x) { return RecFilterDimAndCausality(x, false); }

int RecFilter::max_threads_per_cuda_warp = 0;
int RecFilter::vectorization_width = 0;

void RecFilter::set_max_threads_per_cuda_warp(int v)
{
    if (v % 32 != 0) die("Error: max threads per CUDA warp must be a multiple of 32");     // lib/recfilter.cpp:39-47
    max_threads_per_cuda_warp = v;
}
void RecFilter::set_vectorization_width(int v) { vectorization_width = v; }

// ---------------------------------------------------------------------------------------------
// construction and definition
// ---------------------------------------------------------------------------------------------
RecFilter::RecFilter(string n) : contents(std::make_shared<RecFilterContents>())
{
    static int counter = 0;
    contents->name = n.empty() ? "RecFilter_" + std::to_string(counter++) : n;
}

RecFilter& RecFilter::operator=(const RecFilter& r) { contents = r.contents; return *this; }
string RecFilter::name() const { return contents->name; }

RecFilterRefVar RecFilter::operator()(RecFilterDim x) { return RecFilterRefVar(*this, { x }); }
RecFilterRefVar RecFilter::operator()(RecFilterDim x, RecFilterDim y) { return RecFilterRefVar(*this, { x, y }); }
RecFilterRefVar RecFilter::operator()(RecFilterDim x, RecFilterDim y, RecFilterDim z) { return RecFilterRefVar(*this, { x, y, z }); }
RecFilterRefVar RecFilter::operator()(vector<RecFilterDim> x) { return RecFilterRefVar(*this, x); }

RecFilterRefExpr RecFilter::operator()(Var x) { return RecFilterRefExpr(*this, { Expr(x) }); }
RecFilterRefExpr RecFilter::operator()(Var x, Var y) { return RecFilterRefExpr(*this, { Expr(x), Expr(y) }); }
RecFilterRefExpr RecFilter::operator()(Var x, Var y, Var z) { return RecFilterRefExpr(*this, { Expr(x), Expr(y), Expr(z) }); }
RecFilterRefExpr RecFilter::operator()(vector<Var> x)
{
    vector<Expr> a;
    for (const Var& v : x) a.push_back(Expr(v));
    return RecFilterRefExpr(*this, a);
}
RecFilterRefExpr RecFilter::operator()(Expr x) { return RecFilterRefExpr(*this, { x }); }
RecFilterRefExpr RecFilter::operator()(Expr x, Expr y) { return RecFilterRefExpr(*this, { x, y }); }
RecFilterRefExpr RecFilter::operator()(Expr x, Expr y, Expr z) { return RecFilterRefExpr(*this, { x, y, z }); }
RecFilterRefExpr RecFilter::operator()(vector<Expr> x) { return RecFilterRefExpr(*this, x); }

void RecFilterRefVar::operator=(Expr pure_def) { rf.define(args, { pure_def }); }
void RecFilterRefVar::operator=(const Tuple& pure_def) { rf.define(args, pure_def.as_vector()); }
void RecFilterRefVar::operator=(FuncRefExpr pure_def) { rf.define(args, { Expr(pure_def) }); }
void RecFilterRefVar::operator=(vector<Expr> pure_def) { rf.define(args, pure_def); }
RecFilterRefVar::operator Expr()
{
    vector<Expr> a;
    for (const RecFilterDim& d : args) a.push_back(Expr(d));
    return Expr(FuncRefExpr(rf.handle(), a));
}
Expr RecFilterRefVar::operator[](int i)
{
    vector<Expr> a;
    for (const RecFilterDim& d : args) a.push_back(Expr(d));
    return FuncRefExpr(rf.handle(), a)[i];
}
RecFilterRefExpr::operator Expr() { return Expr(FuncRefExpr(rf.handle(), args)); }
Expr RecFilterRefExpr::operator[](int i) { return FuncRefExpr(rf.handle(), args)[i]; }

void RecFilter::define(vector<RecFilterDim> pure_args, vector<Expr> pure_def)
{
    RecFilterContents& c = *contents;
    if (pure_args.empty() || pure_def.empty()) die("RecFilter " + c.name + ": empty definition");
    c.channels.clear();
    if (pure_def.size() > 1) {
        // Tuple: one channel filter per element (same dimensions, border mode and, later, scans)
        c.dims = pure_args;
        c.rhs = pure_def;
        c.scans.clear();
        c.src_image.reset(); c.src_filter.reset(); c.stencil.clear(); c.side_image.reset(); c.side_filter.reset();
        if (c.plan) { rf_plan_destroy(c.plan); c.plan = nullptr; }
        for (size_t i = 0; i < pure_def.size(); ++i) {
            RecFilter ch(c.name + "_" + std::to_string(i));
            if (c.clamped) ch.set_clamped_image_border();
            ch.define(pure_args, vector<Expr>(1, pure_def[i]));
            c.channels.push_back(ch.handle());
        }
        c.type = c.channels[0]->type;
        for (auto& ch : c.channels)
            if (!(ch->type == c.type)) die("RecFilter " + c.name + ": the elements of a Tuple must have the same type");
        c.defined = true;
        return;
    }
    const Expr& e = pure_def[0];
    if (!e.defined()) die("RecFilter " + c.name + ": undefined expression in the definition");
    c.dims = pure_args;
    c.rhs = pure_def;
    c.scans.clear();
    c.src_image.reset(); c.src_filter.reset();
    if (c.plan) { rf_plan_destroy(c.plan); c.plan = nullptr; }

    vector<LinTap> taps;
    double constant = 0.0;
    linearize(c, e, IndexEnv(), 1.0, taps, constant);
    if (taps.empty()) die("RecFilter " + c.name + ": the definition does not read an image or a filter");
    if (constant != 0.0) die("RecFilter " + c.name + ": constant terms in the definition are not supported");
    // merge equal taps; everything must read ONE filter result and / or ONE image, with the filter's own
    // dimension order.  With both, the filter result is source 0 and the image source 1.
    std::shared_ptr<BufferData> image;
    std::shared_ptr<RecFilterContents> filter, filter2;
    for (const LinTap& t : taps) {
        if (t.filter) {
            if (!filter || filter == t.filter) filter = t.filter;
            else if (!filter2 || filter2 == t.filter) filter2 = t.filter;
            else die("RecFilter " + c.name + ": the definition may read two filters at most");
        } else { if (image && image != t.image) die("RecFilter " + c.name + ": the definition may read one image only"); image = t.image; }
    }
    if (filter2 && image) die("RecFilter " + c.name + ": the definition may read two filters, or one filter and one image");
    vector<LinTap> merged;
    for (const LinTap& t : taps) {
        if (t.idx.size() != c.dims.size()) die("RecFilter " + c.name + ": dimension mismatch in the definition");
        for (size_t i = 0; i < t.idx.size(); ++i)
            if (t.idx[i].dim != (int)i) die("RecFilter " + c.name + ": indices must use the filter's dimensions in order");
        bool found = false;
        for (LinTap& m : merged) {
            bool same = m.image == t.image && m.filter == t.filter;
            for (size_t i = 0; same && i < t.idx.size(); ++i)
                same = m.idx[i].off == t.idx[i].off && m.idx[i].lo == t.idx[i].lo && m.idx[i].hi == t.idx[i].hi;
            if (same) { m.w += t.w; found = true; break; }
        }
        if (!found) merged.push_back(t);
    }
    c.side_image.reset(); c.side_filter.reset();
    if (c.dev_side) { rf_free(c.dev_side); c.dev_side = nullptr; }
    if (image) {
        if (!image->has_data()) die("RecFilter " + c.name + ": the image in the definition has no data");
        for (size_t i = 0; i < c.dims.size(); ++i)
            if (image->extent[i] < c.dims[i].num_pixels())
                die("RecFilter " + c.name + ": the image is smaller than the filter domain");
    }
    if (filter) {
        if (!filter->defined) die("RecFilter " + c.name + ": the filter called in the definition is not defined");
        if (filter->dims.size() != c.dims.size()) die("RecFilter " + c.name + ": dimension mismatch with the called filter");
        c.src_filter = filter;
        c.type = filter->type;
        if (image) {
            if (image->type != filter->type) die("RecFilter " + c.name + ": the image and the filter in the definition differ in type");
            c.side_image = image;
        }
        if (filter2) {
            if (!filter2->defined) die("RecFilter " + c.name + ": the second filter called in the definition is not defined");
            if (filter2->dims.size() != c.dims.size()) die("RecFilter " + c.name + ": dimension mismatch with the second called filter");
            if (!(filter2->type == filter->type)) die("RecFilter " + c.name + ": the two filters in the definition differ in type");
            for (size_t i = 0; i < c.dims.size(); ++i)
                if (filter2->dims[i].num_pixels() != filter->dims[i].num_pixels())
                    die("RecFilter " + c.name + ": the two filters in the definition differ in extent");
            c.side_filter = filter2;
        }
    } else {
        c.src_image = image;
        c.type = image->type;                                       // type of the filter = type of the RHS (lib/recfilter.cpp:197)
        // a type-converting definition, cast<float>(image16(x, y)) (apps/DoG/diff_gauss.cpp:66-73) or image8(x, y) * 0.5f:
        // the RHS is Float(32) over an integer image -> a float filter, the image is converted on upload.  Any other
        // mismatch is refused rather than run in the image's type with truncated coefficients.
        if (e.defined() && !(e.type() == image->type)) {
            if (e.type() == Float(32) && !image->type.is_float()) c.type = Float(32);
            else die("RecFilter " + c.name + ": type-converting definitions other than integer image -> float are not supported");
        }
    }
    // a single unit tap whose indices are the identity inside the domain is the input itself
    c.stencil.clear();
    bool identity = merged.size() == 1 && merged[0].w == 1.0 && !c.side_image && !c.side_filter;
    for (size_t i = 0; identity && i < c.dims.size(); ++i) {
        const IndexMap& m = merged[0].idx[i];
        identity = m.off == 0 && m.lo <= 0 && m.hi >= (long long)c.dims[i].num_pixels() - 1;
    }
    if (!identity) {
        if (merged.size() > RF_MAX_TAPS) die("RecFilter " + c.name + ": too many taps in the definition");
        if (c.type.bytes() != 4) die("RecFilter " + c.name + ": stencil definitions need a 32-bit element type");
        // factor the first weight out: "(a - b - c + d) / area" becomes unit taps and one scale after the sum.
        // Integer filters: weights are ring elements, nothing is factored out, and a weight that is not an exact
        // integer (a scale such as "/ area": Halide's integer division is not a linear scale) is refused instead of
        // being rounded silently.
        const bool integral = !c.type.is_float();
        if (integral)
            for (const LinTap& t : merged)
                if (t.w != std::floor(t.w) || std::fabs(t.w) > 2147483647.0)
                    die("RecFilter " + c.name + ": an integer definition needs integer weights (division and fractional "
                        "scales of an integer filter are not supported: convert to float first)");
        const double w0 = integral ? 1.0 : (merged[0].w != 0.0 ? merged[0].w : 1.0);
        c.stencil_scale = (float)w0;
        for (const LinTap& t : merged) {
            rf_tap rt;
            std::memset(&rt, 0, sizeof(rt));
            rt.weight = (float)(t.w / w0);
            rt.source = ((c.side_image && !t.filter) || (c.side_filter && t.filter == c.side_filter)) ? 1 : 0;
            for (int d = 0; d < RF_MAX_DIMS; ++d) { rt.lo[d] = INT32_MIN; rt.hi[d] = INT32_MAX; }
            for (size_t i = 0; i < t.idx.size(); ++i) {
                rt.offset[i] = (int32_t)t.idx[i].off;
                rt.lo[i] = (int32_t)std::max<long long>(t.idx[i].lo, INT32_MIN);
                rt.hi[i] = (int32_t)std::min<long long>(t.idx[i].hi, INT32_MAX);
            }
            c.stencil.push_back(rt);
        }
    }
    c.defined = true;
}

void RecFilter::set_clamped_image_border()
{
    if (contents->defined) die("Border clamping must be set before defining the filter");           // lib/recfilter.cpp:252-258
    contents->clamped = true;
}

void RecFilter::add_filter(RecFilterDim x, vector<float> coeff) { add_filter(RecFilterDimAndCausality(x, true), coeff); }

void RecFilter::add_filter(RecFilterDimAndCausality x, vector<float> coeff)
{
    RecFilterContents& c = *contents;
    if (!c.defined) die("Cannot add scans to the filter " + c.name + " before it is defined");       // lib/recfilter.cpp:268-272
    if (coeff.size() < 2) die("Cannot add scan without feed forward and feedback coefficients");     // :274-278
    if (c.tiled) die("Cannot add scans to the filter " + c.name + " after it is tiled");
    const int dim = c.dim_index(x.var().name());
    if (dim < 0) die("Variable " + x.var().name() + " is not one of the dimensions of the filter " + c.name);   // :296-300
    if (c.plan) { rf_plan_destroy(c.plan); c.plan = nullptr; }
    c.scans.push_back({ dim, x.causal(), coeff });
}

// ---------------------------------------------------------------------------------------------
// tiling: recorded as a hint.  The reference rewrites the pipeline here (lib/split.cpp:1850-2080);
// the engine's planner tiles every filter itself, with register tiles sized for B200.
// ---------------------------------------------------------------------------------------------
void RecFilter::split(std::map<string, int> dim_tile)
{
    RecFilterContents& c = *contents;
    if (c.tiled) die("Recursive filter " + c.name + " cannot be tiled twice");                       // lib/split.cpp:1851-1854
    if (!c.defined) die("Recursive filter " + c.name + " must be defined before it is tiled");
    for (const auto& kv : dim_tile) {
        const int dim = c.dim_index(kv.first);
        if (dim < 0) die("Variable " + kv.first + " to be tiled is not one of the dimensions of the filter " + c.name);   // :1967-1971
        bool has_scan = false;
        for (const ScanDef& s : c.scans) has_scan = has_scan || s.dim == dim;
        if (!has_scan) die("Dimension " + kv.first + " of the filter " + c.name + " has no scans, it cannot be tiled");    // :1879-1883
        if (kv.second <= 0 || c.dims[dim].num_pixels() % kv.second != 0)
            die("Tile width must be a positive divisor of the image width in dimension " + kv.first);                     // lib/recfilter.h:311
        c.tiles[kv.first] = kv.second;
    }
    c.tiled = true;
    if (c.plan) { rf_plan_destroy(c.plan); c.plan = nullptr; }
}
void RecFilter::split(RecFilterDim x, int tx) { split(std::map<string, int>{ { x.var().name(), tx } }); }
void RecFilter::split(RecFilterDim x, int tx, RecFilterDim y, int ty)
{
    split(std::map<string, int>{ { x.var().name(), tx }, { y.var().name(), ty } });
}
void RecFilter::split(RecFilterDim x, int tx, RecFilterDim y, int ty, RecFilterDim z, int tz)
{
    split(std::map<string, int>{ { x.var().name(), tx }, { y.var().name(), ty }, { z.var().name(), tz } });
}
void RecFilter::split_all_dimensions(int tx)
{
    std::map<string, int> m;
    for (size_t i = 0; i < contents->dims.size(); ++i) {
        bool has_scan = false;
        for (const ScanDef& s : contents->scans) has_scan = has_scan || s.dim == (int)i;
        if (has_scan) m[contents->dims[i].var().name()] = tx;        // dimensions without scans stay untiled (lib/split.cpp:1888-1898)
    }
    split(m);
}

// ---------------------------------------------------------------------------------------------
// cascade / overlap (lib/reorder.cpp)
// ---------------------------------------------------------------------------------------------
vector<RecFilter> RecFilter::cascade(vector<int> a, vector<int> b) { return cascade(vector<vector<int> >{ a, b }); }

vector<RecFilter> RecFilter::cascade(vector<vector<int> > groups)
{
    RecFilterContents& c = *contents;
    if (c.tiled) die("Cascading must be done before the filter " + c.name + " is tiled");
    if (!c.channels.empty()) die("RecFilter " + c.name + ": cascading a Tuple filter is not supported; cascade its channels");
    const int n = (int)c.scans.size();
    vector<int> group_of(n, -1);
    for (size_t g = 0; g < groups.size(); ++g)
        for (int s : groups[g]) {
            if (s < 0 || s >= n) die("Scan index " + std::to_string(s) + " in cascade() is out of range");
            if (group_of[s] >= 0) die("Scan " + std::to_string(s) + " appears twice in cascade()");
            group_of[s] = (int)g;
        }
    for (int s = 0; s < n; ++s)
        if (group_of[s] < 0) die("Scan " + std::to_string(s) + " is missing from cascade()");
    // scans of one dimension with opposite causality do not commute: their order must survive (lib/reorder.cpp:55-78)
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (c.scans[i].dim == c.scans[j].dim && c.scans[i].causal != c.scans[j].causal && group_of[i] > group_of[j])
                die("Cascade would reorder scans " + std::to_string(i) + " and " + std::to_string(j) +
                    " of opposite causality along the same dimension");
    vector<RecFilter> out;
    for (size_t g = 0; g < groups.size(); ++g) {
        RecFilter f(c.name + "_" + std::to_string(g));
        RecFilterContents& fc = *f.contents;
        fc.dims = c.dims; fc.type = c.type; fc.clamped = c.clamped; fc.defined = true;
        if (g == 0) { fc.rhs = c.rhs; fc.src_image = c.src_image; fc.src_filter = c.src_filter; fc.stencil = c.stencil; fc.stencil_scale = c.stencil_scale; fc.side_image = c.side_image; fc.side_filter = c.side_filter; }
        else        { fc.src_filter = out[g - 1].contents; }
        vector<int> ids = groups[g];
        std::sort(ids.begin(), ids.end());                           // add_filter order inside a group
        for (int s : ids) fc.scans.push_back(c.scans[s]);
        out.push_back(f);
    }
    return out;
}

vector<RecFilter> RecFilter::cascade_by_dimension()
{
    vector<vector<int> > groups;
    for (size_t d = 0; d < contents->dims.size(); ++d) {
        vector<int> g;
        for (size_t s = 0; s < contents->scans.size(); ++s) if (contents->scans[s].dim == (int)d) g.push_back((int)s);
        if (!g.empty()) groups.push_back(g);
    }
    return cascade(groups);
}

vector<RecFilter> RecFilter::cascade_by_causality()
{
    vector<int> causal, anticausal;
    for (size_t s = 0; s < contents->scans.size(); ++s) (contents->scans[s].causal ? causal : anticausal).push_back((int)s);
    vector<vector<int> > groups;
    if (!causal.empty()) groups.push_back(causal);
    if (!anticausal.empty()) groups.push_back(anticausal);
    return cascade(groups);
}

RecFilter RecFilter::overlap_to_higher_order_filter(RecFilter fB, string overlap_name)
{
    RecFilterContents& a = *contents;
    RecFilterContents& b = *fB.contents;
    if (a.tiled || b.tiled) die("Overlapping directive overlap() cannot be used after the filter is already tiled");
    if (!a.channels.empty() || !b.channels.empty()) die("Overlapping Tuple filters is not supported; overlap their channels");
    if (b.src_filter.get() != &a)
        die("Filters cannot be overlapped because the input to second does not match the output of the first");
    if (a.clamped != b.clamped) die("Filters cannot be overlapped because one clamps image border while the other does not");
    if (a.type != b.type) die("Filters cannot be overlapped because they have different types");
    RecFilter ab(overlap_name);
    RecFilterContents& c = *ab.contents;
    c.dims = a.dims; c.type = a.type; c.clamped = a.clamped; c.defined = true;
    c.rhs = a.rhs; c.src_image = a.src_image; c.src_filter = a.src_filter; c.stencil = a.stencil; c.stencil_scale = a.stencil_scale; c.side_image = a.side_image; c.side_filter = a.side_filter;
    for (size_t d = 0; d < a.dims.size(); ++d) {
        vector<const ScanDef*> sa, sb;
        for (const ScanDef& s : a.scans) if (s.dim == (int)d) sa.push_back(&s);
        for (const ScanDef& s : b.scans) if (s.dim == (int)d) sb.push_back(&s);
        if (sa.size() != sb.size())
            die("Filters cannot be overlapped because they have different num scans in dimension " + std::to_string(d));
        for (size_t j = 0; j < sa.size(); ++j) {
            if (sa[j]->causal != sb[j]->causal)
                die("Filters cannot be overlapped because they have different causality in scan " + std::to_string(j) +
                    " of dimension " + std::to_string(d));
            vector<float> fa(sa[j]->coeff.begin() + 1, sa[j]->coeff.end());
            vector<float> fb(sb[j]->coeff.begin() + 1, sb[j]->coeff.end());
            vector<float> coeff = overlap_feedback_coeff(fa, fb);
            coeff.insert(coeff.begin(), sa[j]->coeff[0] * sb[j]->coeff[0]);
            c.scans.push_back({ (int)d, sa[j]->causal, coeff });
        }
    }
    return ab;
}

// ---------------------------------------------------------------------------------------------
// execution
// ---------------------------------------------------------------------------------------------
Target RecFilter::target() { return Target(); }
void RecFilter::apply_bounds() {}
void RecFilter::compile_jit(string) { if (contents->defined) build_plan(*contents); }

Realization RecFilter::realize()
{
    RecFilterContents& c = *contents;
    if (!c.defined) die("Filter " + c.name + " cannot be realized before it is defined");
    if (!c.channels.empty()) {                                       // Tuple: one Buffer per element
        sync_channels(c);
        vector<Buffer> bufs;
        for (auto& ch : c.channels) {
            void* d = evaluate_device(*ch);
            vector<int> e;
            for (const RecFilterDim& dim : c.dims) e.push_back(dim.num_pixels());
            Buffer b(ch->type, e);
            engine_check(rf_memcpy_d2h(b.host_ptr(), d, b.size_in_bytes()), "rf_memcpy_d2h");
            bufs.push_back(b);
        }
        return Realization(bufs);
    }
    vector<int> ext;
    for (const RecFilterDim& d : c.dims) ext.push_back(d.num_pixels());
    Buffer out(c.type, ext);
    if (rf_mgpu* m = mgpu_for(c)) {
        if (rf_mgpu_execute_host(m, unit_first(c).src_image->host_data(), out.host_ptr()) != RF_OK)
            die(string("rf_mgpu_execute_host failed: ") + rf_mgpu_last_error());
    } else {
        void* dev = evaluate_device(c);
        engine_check(rf_memcpy_d2h(out.host_ptr(), dev, out.size_in_bytes()), "rf_memcpy_d2h");
    }
    if (const char* path = getenv("RECFILTER_DUMP_FILTER")) {
        // one JSON line per realize(): the filter chain as declared (used to label golden vectors)
        vector<RecFilterContents*> chain;
        collect_chain(c, chain);
        std::ofstream f(path, std::ios::app);
        f << "{\"name\": \"" << c.name << "\", \"extent\": [";
        for (size_t i = 0; i < ext.size(); ++i) f << (i ? ", " : "") << ext[i];
        f << "], \"dtype\": \"" << (c.type.is_float() ? "float" : (c.type.is_int() ? "int" : "uint")) << c.type.bits
          << "\", \"border\": \"" << (c.clamped ? "clamp" : "zero") << "\", \"stages\": [";
        for (size_t g = 0; g < chain.size(); ++g) {
            f << (g ? ", " : "") << "[";
            for (size_t k = 0; k < chain[g]->scans.size(); ++k) {
                const ScanDef& sc = chain[g]->scans[k];
                f << (k ? ", " : "") << "[" << sc.dim << ", " << (sc.causal ? "true" : "false") << ", [";
                f << std::setprecision(9);
                for (size_t j = 0; j < sc.coeff.size(); ++j) f << (j ? ", " : "") << sc.coeff[j];
                f << "]]";
            }
            f << "]";
        }
        f << "]}\n";
    }
    return Realization(vector<Buffer>{ out });
}

float RecFilter::profile(int iterations)
{
    RecFilterContents& c = *contents;
    if (!c.defined) die("Filter " + c.name + " cannot be profiled before it is defined");
    if (iterations < 1) iterations = 1;
    if (!c.channels.empty()) {                                       // Tuple: the channels run one after the other
        sync_channels(c);
        float total = 0.0f;
        for (auto& ch : c.channels) { RecFilter f; f.contents = ch; total += f.profile(iterations); }
        cerr << c.name << ": " << total << " ms per iteration for " << c.channels.size() << " Tuple elements" << endl;
        return total;
    }
    if (rf_mgpu* m = mgpu_for(c)) {
        float per_iter = 0.0f;
        if (rf_mgpu_profile(m, unit_first(c).src_image->host_data(), iterations, &per_iter) != RF_OK)
            die(string("rf_mgpu_profile failed: ") + rf_mgpu_last_error());
        cerr << c.name << ": " << per_iter << " ms per iteration over " << iterations << " iteration(s) on " << rf_mgpu_ngpus(m)
             << " GPUs, " << (double(c.count()) / (per_iter * 1e-3) / 1e9) << " Gsamples/s" << endl;
        return per_iter;
    }
    vector<std::pair<RecFilterContents*, RecFilterContents*>> units;
    bool owned = false;
    void* root_in = nullptr;
    EpilogueMatch epi;
    bool fused_epilogue = epilogue_match(c, epi);
    if (fused_epilogue) {                                             // a * image + b * filter(image): one plan, epilogue in its store
        root_in = upload_image(c, *c.side_image); owned = true;
        if (!run_epilogue_unit(c, epi, root_in)) {
            fused_epilogue = false;
            engine_check(rf_synchronize(), "rf_synchronize"); engine_check(rf_free(root_in), "rf_free");
        }
    }
    if (!fused_epilogue) {
        collect_units(c, units);
        root_in = input_device(*units[0].first, owned);               // the image stays resident: kernels only are timed
    }
    bool warm = false;
    auto run_chain = [&]() {
        if (fused_epilogue) { run_epilogue_unit(c, epi, root_in); return; }
        const void* in = root_in;
        for (auto& u : units) in = run_unit(*u.first, *u.second, in, !warm);
    };
    g_resident_inputs = true;                                         // (side chains evaluated inside run_unit upload once)
    run_chain();                                                      // warm-up (lib/recfilter.cpp:995-997)
    warm = true;
    void* clock = nullptr;
    engine_check(rf_clock_begin(nullptr, &clock), "rf_clock_begin");
    for (int i = 0; i < iterations; ++i) run_chain();
    float ms = 0.0f;
    engine_check(rf_clock_end(clock, nullptr, &ms), "rf_clock_end");
    release_resident_inputs();
    if (owned) engine_check(rf_free(root_in), "rf_free");
    const float per_iter = ms / float(iterations);
    cerr << c.name << ": " << per_iter << " ms per iteration over " << iterations << " iteration(s), "
         << (double(c.count()) / (per_iter * 1e-3) / 1e9) << " Gsamples/s" << endl;
    return per_iter;
}

Func RecFilter::as_func() { return Func(contents, contents->name); }
Func RecFilter::func(string func_name) { return Func(contents, func_name); }

// ---------------------------------------------------------------------------------------------
// schedules: accepted, ignored (the planner replaces lib/schedule.cpp)
// ---------------------------------------------------------------------------------------------
RecFilterSchedule RecFilter::intra_schedule(int) { return RecFilterSchedule(*this, { contents->name }); }
RecFilterSchedule RecFilter::inter_schedule() { return RecFilterSchedule(*this, { contents->name }); }
RecFilterSchedule RecFilter::full_schedule() { return RecFilterSchedule(*this, { contents->name }); }
// compute_at merges a filter into its pointwise consumer (lib/recfilter.cpp:473-573).  Nothing to record here: the merge
// the in-scope apps ask for (apps/usm: blur into the unsharp mask) is found from the definitions themselves and done by
// the store of the filter's last kernel (epilogue_match / rf_options.epilogue), whether or not compute_at was called.
void RecFilter::compute_at(RecFilter) {}
void RecFilter::compute_at(Func, Var) {}
void RecFilter::gpu_auto_full_schedule(int) {}
void RecFilter::gpu_auto_schedule(int)
{
    if (max_threads_per_cuda_warp <= 0)
        die("Use RecFilter::set_max_threads_per_cuda_warp() to specify the maximum number of threads in each CUDA warp");   // lib/recfilter.cpp:699-704
}
void RecFilter::gpu_auto_inter_schedule() {}
void RecFilter::gpu_auto_intra_schedule(int) {}
void RecFilter::cpu_auto_schedule() {}
void RecFilter::cpu_auto_full_schedule() {}
void RecFilter::cpu_auto_inter_schedule() {}
void RecFilter::cpu_auto_intra_schedule() {}

VarTag RecFilter::full(int i) { return VarTag(FULL, i); }
VarTag RecFilter::inner(int i) { return VarTag(INNER, i); }
VarTag RecFilter::outer(int i) { return VarTag(OUTER, i); }
VarTag RecFilter::tail() { return VarTag(TAIL); }
VarTag RecFilter::full_scan() { return VarTag(SCAN, 0); }
VarTag RecFilter::inner_scan() { return VarTag(SCAN, 1); }
VarTag RecFilter::outer_scan() { return VarTag(SCAN, 2); }
VarTag RecFilter::inner_channels() { return VarTag(CHANNEL, 0); }
VarTag RecFilter::outer_channels() { return VarTag(CHANNEL, 1); }

// ---------------------------------------------------------------------------------------------
// printing: the filter as declared, and the launch plan when a device is present
// ---------------------------------------------------------------------------------------------
string RecFilter::print_synopsis() const
{
    const RecFilterContents& c = *contents;
    std::ostringstream s;
    s << "RecFilter " << c.name << "(";
    for (size_t i = 0; i < c.dims.size(); ++i) s << (i ? ", " : "") << c.dims[i].var().name() << ":" << c.dims[i].num_pixels();
    s << ")" << (c.clamped ? " clamped border" : " zero border");
    if (c.src_filter) s << ", input = " << c.src_filter->name;
    s << "\n";
    for (size_t i = 0; i < c.scans.size(); ++i) {
        s << "  scan " << i << ": " << (c.scans[i].causal ? "+" : "-") << c.dims[c.scans[i].dim].var().name() << " {";
        for (size_t k = 0; k < c.scans[i].coeff.size(); ++k) s << (k ? ", " : "") << c.scans[i].coeff[k];
        s << "}\n";
    }
    for (const auto& kv : c.tiles) s << "  split " << kv.first << " by " << kv.second << " (hint)\n";
    return s.str();
}
string RecFilter::print_functions() const { return print_synopsis(); }
string RecFilter::print_schedule() const
{
    if (contents->defined && rf_device_count() > 0) {
        build_plan(*contents);
        char buf[8192];
        if (rf_plan_describe(contents->plan, buf, sizeof(buf)) == RF_OK) return string(buf);
    }
    return "launch plan: built at the first realize() on a CUDA device\n";
}
string RecFilter::print_hl_code() const { return print_synopsis(); }

std::ostream& operator<<(std::ostream& s, const RecFilter& r)
{
    s << r.print_synopsis() << r.print_schedule();
    return s;
}
std::ostream& operator<<(std::ostream& s, const RecFilterDim& f)
{
    s << f.var().name() << "[" << f.num_pixels() << "]";
    return s;
}
std::ostream& operator<<(std::ostream& s, const Func& f)
{
    s << "Func " << f.name();
    return s;
}

// ---------------------------------------------------------------------------------------------
// command line of the tests and apps (lib/recfilter_utils.cpp:31-112): same flags, same defaults
// ---------------------------------------------------------------------------------------------
Arguments::Arguments(int argc, char** argv)
    : width(4096), max_width(4096), min_width(4096), block(32), iterations(1), nocheck(false), noschedule(false)
{
    const string usage = string("\nUsage\n ") + argv[0] +
        " [-width|-w w] [-tile|-t b] [-iter i] [-nocheck] [-help]\n\n"
        "\twidth    image width, 0 sweeps all widths and forces -nocheck [default = 4096]\n"
        "\ttile     tile width for splitting each dimension [default = 32]\n"
        "\tnocheck  do not check against the reference solution, forced when width=0 or iter>1 [default = false]\n"
        "\titer     number of profiling iterations [default = 1]\n"
        "\thelp     show this message\n";
    auto is = [](const string& o, const char* n) { return o == string("-") + n || o == string("--") + n; };
    try {
        for (int i = 1; i < argc; ++i) {
            const string o = argv[i];
            auto value = [&](const char* what) {
                if (i + 1 >= argc) throw std::runtime_error(string(what) + " requires an integer value");
                return atoi(argv[++i]);
            };
            if (is(o, "help")) throw std::runtime_error("Showing help message");
            else if (is(o, "nocheck")) nocheck = true;
            else if (is(o, "iter")) iterations = value("-iter");
            else if (is(o, "w") || is(o, "width")) width = value("-width");
            else if (is(o, "t") || is(o, "tile")) block = value("-tile");
            else throw std::runtime_error("Bad command line option " + o);
        }
        if (block <= 0 || width % block) throw std::runtime_error("Width should be a multiple of block size");
        if (width) { max_width = width; min_width = width; }
        else { min_width = 2 * block; max_width = (4096 / block) * block; nocheck = true; }
        if (iterations > 1) nocheck = true;
    } catch (std::runtime_error& e) {
        cerr << endl << e.what() << endl << usage << endl;
        exit(EXIT_FAILURE);
    }
}
