/*
 * recfilter.h -- the RecFilter C++ operator surface, implemented on the B200 engine.
 *
 * Same class / method names, argument meaning and error behaviour as the reference's
 * public header (/root/reference/lib/recfilter.h:68-855) so that its tests and apps compile
 * and run unchanged:
 *
 *     RecFilterDim x("x", w), y("y", h);
 *     RecFilter F;  F(x,y) = image(x,y);
 *     F.add_filter(+x, {b0, a1, .., ar});  F.add_filter(-y, ...);   // lib/recfilter.cpp:260-392
 *     F.split(x, tile, y, tile);                                   // lib/split.cpp:1850-2111
 *     Halide::Realization out = F.realize();                       // lib/recfilter.cpp:984-989
 *
 * What differs underneath: there is no Halide pipeline.  define/add_filter/split only
 * record the filter; realize()/profile() hand the scan list to the launch planner behind the
 * C ABI of include/recfilter_b200.h, which runs hand-written sm_100a kernels.  The scheduling
 * handles (intra_schedule(), gpu_auto_schedule(), RecFilterSchedule::*) are accepted and
 * ignored: mapping work to the GPU is the planner's job (lib/schedule.cpp is replaced, not
 * ported).  Errors follow the reference's convention: message on stderr, then assert(false).
 */
#ifndef _RECURSIVE_FILTER_H_
#define _RECURSIVE_FILTER_H_

#include <algorithm>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <Halide.h>

class RecFilter;
class RecFilterSchedule;
class RecFilterRefVar;
class RecFilterRefExpr;
struct RecFilterContents;

/** Filter dimension: variable name + extent of the image along it (lib/recfilter.h:68-95) */
class RecFilterDim {
public:
    RecFilterDim() : e(0) {}
    RecFilterDim(std::string var_name, int var_extent) : v(var_name), e(var_extent) {}
    Halide::Var var() const { return v; }
    int num_pixels() const { return e; }
    operator Halide::Expr() const { return Halide::Internal::Variable::make(Halide::Int(32), v.name()); }
private:
    Halide::Var v;
    int e;
};

/** Filter dimension + causality (lib/recfilter.h:98-128) */
class RecFilterDimAndCausality {
public:
    RecFilterDimAndCausality() : c(true) {}
    RecFilterDimAndCausality(RecFilterDim rec_var, bool causal) : r(rec_var), c(causal) {}
    Halide::Var var() const { return r.var(); }
    int num_pixels() const { return r.num_pixels(); }
    bool causal() const { return c; }
    operator Halide::Expr() const { return Halide::Internal::Variable::make(Halide::Int(32), r.var().name()); }
private:
    RecFilterDim r;
    bool c;
};

/** +x: causal scan, -x: anticausal scan (lib/recfilter.h:135-139) */
RecFilterDimAndCausality operator+(RecFilterDim x);
RecFilterDimAndCausality operator-(RecFilterDim x);

/** Tags naming groups of loop variables in the reference's schedules (lib/recfilter.h:635-656).
 * Kept so that hand-written schedules compile; they carry no meaning for the planner. */
enum VariableTag { INVALID = 0, FULL, INNER, OUTER, TAIL, SCAN, CHANNEL, SPLIT };
class VarTag {
public:
    VarTag() : tag(INVALID), cnt(-1) {}
    VarTag(const VarTag& t) : tag(t.tag), cnt(t.cnt) {}
    VarTag(const VariableTag& t) : tag(t), cnt(-1) {}
    VarTag(const VarTag& t, int i) : tag(t.tag), cnt(i) {}
    VarTag(const VariableTag& t, int i) : tag(t), cnt(i) {}
    VarTag(int i) : tag(INVALID), cnt(i) {}
    VarTag& operator=(const VarTag& t) { tag = t.tag; cnt = t.cnt; return *this; }
    VarTag& operator=(const VariableTag& t) { tag = t; cnt = -1; return *this; }
    int as_integer() const { return (int)tag * 64 + cnt + 1; }
    VarTag split_var() const { return VarTag(SPLIT, cnt); }
    int count() const { return cnt; }
    bool has_count() const { return cnt >= 0; }
    int check(const VariableTag& t) const { return tag == t; }
    bool same_except_count(const VarTag& t) const { return tag == t.tag; }
private:
    VariableTag tag;
    int cnt;
};

/** Recursive filter: a ref-counted handle (copies alias, lib/recfilter.cpp:141-144) */
class RecFilter {
public:
    RecFilter(std::string name = "");
    RecFilter& operator=(const RecFilter& r);
    std::string name() const;

    /** left-hand side of a definition: F(x,y) = ... */
    RecFilterRefVar operator()(RecFilterDim x);
    RecFilterRefVar operator()(RecFilterDim x, RecFilterDim y);
    RecFilterRefVar operator()(RecFilterDim x, RecFilterDim y, RecFilterDim z);
    RecFilterRefVar operator()(std::vector<RecFilterDim> x);
    /** the filter's result as an expression of another pipeline stage */
    RecFilterRefExpr operator()(Halide::Var x);
    RecFilterRefExpr operator()(Halide::Var x, Halide::Var y);
    RecFilterRefExpr operator()(Halide::Var x, Halide::Var y, Halide::Var z);
    RecFilterRefExpr operator()(std::vector<Halide::Var> x);
    RecFilterRefExpr operator()(Halide::Expr x);
    RecFilterRefExpr operator()(Halide::Expr x, Halide::Expr y);
    RecFilterRefExpr operator()(Halide::Expr x, Halide::Expr y, Halide::Expr z);
    RecFilterRefExpr operator()(std::vector<Halide::Expr> x);

    void define(std::vector<RecFilterDim> pure_args, std::vector<Halide::Expr> pure_def);

    Halide::Target target();
    void apply_bounds();
    /** builds the launch plan now instead of at the first realize() (the reference JIT-compiles here) */
    void compile_jit(std::string filename = "");
    /** run the filter; host buffers in, host buffers out */
    Halide::Realization realize();
    /** milliseconds per iteration, device resident data, CUDA events (lib/recfilter.cpp:991-1016) */
    float profile(int iterations);

    void add_filter(RecFilterDim x, std::vector<float> coeff);
    void add_filter(RecFilterDimAndCausality x, std::vector<float> coeff);
    void set_clamped_image_border();

    Halide::Func as_func();
    Halide::Func func(std::string func_name);

    void split_all_dimensions(int tx);
    void split(RecFilterDim x, int tx);
    void split(RecFilterDim x, int tx, RecFilterDim y, int ty);
    void split(RecFilterDim x, int tx, RecFilterDim y, int ty, RecFilterDim z, int tz);
    void split(std::map<std::string, int> dims);

    std::vector<RecFilter> cascade(std::vector<int> a, std::vector<int> b);
    std::vector<RecFilter> cascade(std::vector<std::vector<int> > scan);
    std::vector<RecFilter> cascade_by_causality();
    std::vector<RecFilter> cascade_by_dimension();
    RecFilter overlap_to_higher_order_filter(RecFilter fA, std::string name = "O");

    RecFilterSchedule intra_schedule(int id = 0);
    RecFilterSchedule inter_schedule();
    RecFilterSchedule full_schedule();
    void compute_at(RecFilter external);
    void compute_at(Halide::Func external, Halide::Var granularity);
    void gpu_auto_full_schedule(int tile_width = 32);
    void gpu_auto_schedule(int tile_width = 32);
    void gpu_auto_inter_schedule();
    void gpu_auto_intra_schedule(int id);
    void cpu_auto_schedule();
    void cpu_auto_full_schedule();
    void cpu_auto_inter_schedule();
    void cpu_auto_intra_schedule();

    VarTag full(int i = -1);
    VarTag inner(int i = -1);
    VarTag outer(int i = -1);
    VarTag tail();
    VarTag full_scan();
    VarTag inner_scan();
    VarTag outer_scan();
    VarTag inner_channels();
    VarTag outer_channels();

    std::string print_functions() const;
    std::string print_synopsis() const;
    std::string print_schedule() const;
    std::string print_hl_code() const;

    static void set_max_threads_per_cuda_warp(int v);
    static void set_vectorization_width(int v);

    /** engine-side handle (not part of the reference surface) */
    std::shared_ptr<RecFilterContents> handle() const { return contents; }
    explicit RecFilter(std::shared_ptr<RecFilterContents> c) : contents(c) {}

protected:
    friend class RecFilterSchedule;
    std::shared_ptr<RecFilterContents> contents;
    static int max_threads_per_cuda_warp;
    static int vectorization_width;
};

/** Chainable schedule handle (lib/recfilter.h:516-566): every directive is accepted and ignored. */
class RecFilterSchedule {
public:
    RecFilterSchedule(RecFilter& r, std::vector<std::string> fl) : recfilter(r), func_list(fl) {}
    RecFilterSchedule& compute_globally() { return *this; }
    RecFilterSchedule& compute_locally() { return *this; }
    RecFilterSchedule& fuse(VarTag, VarTag) { return *this; }
    RecFilterSchedule& split(VarTag, int) { return *this; }
    RecFilterSchedule& split(VarTag, int, VarTag) { return *this; }
    RecFilterSchedule& split(VarTag, int, VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder(std::vector<VarTag>) { return *this; }
    RecFilterSchedule& reorder(VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder(VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder(VarTag, VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder(VarTag, VarTag, VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder(VarTag, VarTag, VarTag, VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder(VarTag, VarTag, VarTag, VarTag, VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& storage_layout(VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder_storage(std::vector<VarTag>) { return *this; }
    RecFilterSchedule& reorder_storage(VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder_storage(VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder_storage(VarTag, VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& reorder_storage(VarTag, VarTag, VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& unroll(VarTag, int = 0) { return *this; }
    RecFilterSchedule& parallel(VarTag, int = 0) { return *this; }
    RecFilterSchedule& vectorize(VarTag, int = 0) { return *this; }
    RecFilterSchedule& gpu_threads(VarTag) { return *this; }
    RecFilterSchedule& gpu_threads(VarTag, VarTag) { return *this; }
    RecFilterSchedule& gpu_threads(VarTag, VarTag, VarTag) { return *this; }
    RecFilterSchedule& gpu_blocks(VarTag) { return *this; }
    RecFilterSchedule& gpu_blocks(VarTag, VarTag) { return *this; }
    RecFilterSchedule& gpu_blocks(VarTag, VarTag, VarTag) { return *this; }
protected:
    bool empty() { return func_list.empty(); }
    friend class RecFilter;
private:
    RecFilter recfilter;
    std::vector<std::string> func_list;
};

/** F(x,y) on the left of '=' (lib/recfilter.h:580-608) */
class RecFilterRefVar {
public:
    RecFilterRefVar(RecFilter r, std::vector<RecFilterDim> a) : rf(r), args(a) {}
    void operator=(Halide::Expr pure_def);
    void operator=(const Halide::Tuple& pure_def);
    void operator=(Halide::FuncRefExpr pure_def);
    void operator=(std::vector<Halide::Expr> pure_def);
    operator Halide::Expr();
    Halide::Expr operator[](int);
private:
    RecFilter rf;
    std::vector<RecFilterDim> args;
};

/** F(x,y) used as a value (lib/recfilter.h:611-627) */
class RecFilterRefExpr {
public:
    RecFilterRefExpr(RecFilter r, std::vector<Halide::Expr> a) : rf(r), args(a) {}
    operator Halide::Expr();
    Halide::Expr operator[](int);
private:
    RecFilter rf;
    std::vector<Halide::Expr> args;
};

std::ostream& operator<<(std::ostream& s, const RecFilter& r);
std::ostream& operator<<(std::ostream& s, const RecFilterDim& f);
std::ostream& operator<<(std::ostream& s, const Halide::Func& f);

/** Command line arguments of the tests and apps (lib/recfilter.h:672-684) */
class Arguments {
public:
    int width;        ///< image width
    int max_width;    ///< max image width
    int min_width;    ///< min image width
    int block;        ///< tile width
    int iterations;   ///< profiling iterations
    bool nocheck;     ///< skip the check against the reference solution
    bool noschedule;  ///< never set by the parser (as in the reference)
    Arguments(int argc, char** argv);
};

// ---------------------------------------------------------------------------------------------
// harness templates used by every test / app (lib/recfilter.h:691-855)
// ---------------------------------------------------------------------------------------------

/** "Random" image: the reference draws from [MIN, MAX] with MIN == MAX == 1, i.e. all ones
 * (lib/recfilter.h:695-696,710); kept, so that the programs' own checks see the same input. */
template <typename T>
Halide::Image<T> generate_random_image(size_t w, size_t h = 0, size_t c = 0, size_t d = 0)
{
    const int lo = 1, span = 1;
    Halide::Image<T> image;
    size_t n = 0;
    if (w && h && c && d) { image = Halide::Image<T>((int)w, (int)h, (int)c, (int)d); n = w * h * c * d; }
    else if (w && h && c) { image = Halide::Image<T>((int)w, (int)h, (int)c); n = w * h * c; }
    else if (w && h)      { image = Halide::Image<T>((int)w, (int)h); n = w * h; }
    else if (w)           { image = Halide::Image<T>((int)w); n = w; }
    T* p = image.data();
    for (size_t i = 0; i < n; ++i) p[i] = T(lo + (rand() % span));
    return image;
}

/** Print an image, x fastest, planes separated by "--" (lib/recfilter.h:745-788) */
template <typename T>
std::ostream& operator<<(std::ostream& s, Halide::Image<T> image)
{
    const int width = 4;
    const int nd = image.dimensions();
    const int ex = nd > 0 ? image.extent(0) : 0, ey = nd > 1 ? image.extent(1) : 1;
    const int ez = nd > 2 ? image.extent(2) : 1, ew = nd > 3 ? image.extent(3) : 1;
    for (int w = 0; w < ew; ++w) {
        for (int z = 0; z < ez; ++z) {
            for (int y = 0; y < ey; ++y) {
                for (int x = 0; x < ex; ++x) {
                    if (nd == 1) s << std::setw(width) << image(x) << " ";
                    else         s << std::setw(width) << float(image(x, y, z, w)) << " ";
                }
                s << "\n";
            }
            if (nd > 2) s << "--\n";
        }
        if (nd > 3) s << "--\n";
    }
    return s;
}

/** Relative error report: 100 * |ref - out| / (ref + 1e-9) per sample, max and mean
 * (lib/recfilter.h:793-826) */
template <typename T>
class CheckResult {
public:
    float max_diff;              ///< max relative error in percent
    float mean_diff;             ///< mean relative error in percent
    Halide::Image<T> ref;        ///< reference solution
    Halide::Image<T> out;        ///< engine output
    Halide::Image<float> diff;   ///< per sample difference
    CheckResult(Halide::Image<T> r, Halide::Image<T> o) : max_diff(0.0f), mean_diff(0.0f), ref(r), out(o)
    {
        assert(r.width() == o.width());
        assert(r.height() == o.height());
        assert(r.channels() == o.channels());
        const int W = r.width(), H = r.height(), C = r.channels();
        diff = Halide::Image<float>(W, H, C);
        double sum = 0.0;
        for (int z = 0; z < C; ++z)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const float dv = float(r(x, y, z)) - float(o(x, y, z));
                    diff(x, y, z) = dv;
                    const float re = float(100.0 * std::abs(dv) / (double(r(x, y, z)) + 1e-9));
                    sum += re;
                    max_diff = std::max(re, max_diff);
                }
        mean_diff = float(sum / (double(W) * H * C));
    }
};

template <typename T>
class CheckResultVerbose : public CheckResult<T> {
public:
    CheckResultVerbose(Halide::Image<T> r, Halide::Image<T> o) : CheckResult<T>(r, o) {}
};

template <typename T>
std::ostream& operator<<(std::ostream& s, const CheckResult<T>& v)
{
    s << "Max  relative error = " << v.max_diff << " % \n";
    s << "Mean relative error = " << v.mean_diff << " % \n\n";
    return s;
}

template <typename T>
std::ostream& operator<<(std::ostream& s, const CheckResultVerbose<T>& v)
{
    s << "Reference" << "\n" << v.ref << "\n";
    s << "Halide output" << "\n" << v.out << "\n";
    s << "Difference " << "\n" << v.diff << "\n";
    s << "Max  relative error = " << v.max_diff << " % \n";
    s << "Mean relative error = " << v.mean_diff << " % \n\n";
    return s;
}

#endif // _RECURSIVE_FILTER_H_
