// cuda_sorted_set.hpp — CudaSortedSet: the reference's Set concept over a device-resident sorted set.
//
// Member names, argument meaning and move-only semantics follow SortedSetBase<int32_t>
// (gms/representations/sets/sorted_set.h:22-270), so an algorithm written against the Set concept — Tomita's pivot
// rule `cand.intersect_count(graph.out_neigh(v))` (maximal_clique_enum/sequential/tomita.h:17-31), the pull step of the
// approximate degeneracy order (preprocessing/parallel/degeneracy_approx_set.h:77) — compiles against this class and
// runs on the GPU.  Every set operation is a kernel over set handles (include/gmsb.h: gmsb_set_*); the result of
// intersect / union_with / difference is a new set in HBM.  There is no host implementation of the set algebra here:
// without a CUDA device every operation throws (GMSB_ERR_CUDA).
//
// Iteration (begin / end) is over a host copy fetched on first use and dropped by any modifying operation — it is there
// for callers that walk a small result (a clique, a pivot's candidates), not for the hot path.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <vector>

#include "cuda_set_graph.hpp"

namespace gms_b200 {

class CudaSortedSet {
public:
    using SetElement = NodeId;

    CudaSortedSet() { check(gmsb_set_from_host(nullptr, 0, &h_)); }
    explicit CudaSortedSet(gmsb_set_t h) : h_(h) {}
    // SortedSetBase(const SetElement*, size_t): the elements are sorted on construction (sorted_set.h:64-66)
    CudaSortedSet(const SetElement *start, size_t count) { check(gmsb_set_from_host(start, (int64_t)count, &h_)); }
    explicit CudaSortedSet(const std::vector<SetElement> &v) : CudaSortedSet(v.data(), v.size()) {}
    CudaSortedSet(std::initializer_list<SetElement> l) : CudaSortedSet(std::vector<SetElement>(l)) {}
    CudaSortedSet(SetElement singleton) : CudaSortedSet(&singleton, 1) {}                 // sorted_set.h:55-56
    CudaSortedSet(CudaSortedSet &&o) noexcept : h_(o.h_), host_(std::move(o.host_)), fetched_(o.fetched_) { o.h_ = nullptr; }
    CudaSortedSet &operator=(CudaSortedSet &&o) noexcept {
        if (this != &o) { release(); h_ = o.h_; host_ = std::move(o.host_); fetched_ = o.fetched_; o.h_ = nullptr; }
        return *this;
    }
    CudaSortedSet(const CudaSortedSet &) = delete;                                        // move-only, sorted_set.h:33-39
    CudaSortedSet &operator=(const CudaSortedSet &) = delete;
    ~CudaSortedSet() { release(); }

    static CudaSortedSet Range(int bound) {                                               // sorted_set.h:246-251
        gmsb_set_t h = nullptr;
        check(gmsb_set_range(bound, &h));
        return CudaSortedSet(h);
    }
    // SetGraph::out_neigh(v) as a set: a view of the graph's CSR in HBM (the graph must outlive it)
    static CudaSortedSet Neighbourhood(const CudaSetGraph &g, NodeId v) {
        gmsb_set_t h = nullptr;
        check(gmsb_set_neighbourhood(g.handle(), v, &h));
        return CudaSortedSet(h);
    }

    CudaSortedSet clone() const { gmsb_set_t h = nullptr; check(gmsb_set_clone(h_, &h)); return CudaSortedSet(h); }
    size_t cardinality() const { int64_t n = 0; check(gmsb_set_cardinality(h_, &n)); return (size_t)n; }
    bool contains(SetElement x) const { int f = 0; check(gmsb_set_contains(h_, x, &f)); return f != 0; }
    void toArray(SetElement *array) const { check(gmsb_set_to_host(h_, array)); }       // sorted_set.h:258-262

    CudaSortedSet intersect(const CudaSortedSet &o) const { return op(GMSB_SET_INTERSECT, o); }
    void intersect_inplace(const CudaSortedSet &o) { op_inplace(GMSB_SET_INTERSECT, o); }
    size_t intersect_count(const CudaSortedSet &o) const { return op_count(GMSB_SET_INTERSECT, o); }
    CudaSortedSet union_with(const CudaSortedSet &o) const { return op(GMSB_SET_UNION, o); }
    CudaSortedSet union_with(SetElement x) const { CudaSortedSet r = clone(); r.add(x); return r; }
    void union_inplace(const CudaSortedSet &o) { op_inplace(GMSB_SET_UNION, o); }
    void union_inplace(SetElement x) { add(x); }
    size_t union_count(const CudaSortedSet &o) const { return op_count(GMSB_SET_UNION, o); }
    CudaSortedSet difference(const CudaSortedSet &o) const { return op(GMSB_SET_DIFFERENCE, o); }
    CudaSortedSet difference(SetElement x) const { CudaSortedSet r = clone(); r.remove(x); return r; }
    void difference_inplace(const CudaSortedSet &o) { op_inplace(GMSB_SET_DIFFERENCE, o); }
    void difference_inplace(SetElement x) { remove(x); }
    void add(SetElement x) { check(gmsb_set_add(h_, x)); fetched_ = false; }
    void remove(SetElement x) { check(gmsb_set_remove(h_, x)); fetched_ = false; }

    bool operator==(const CudaSortedSet &o) const { int f = 0; check(gmsb_set_equal(h_, o.h_, &f)); return f != 0; }
    bool operator!=(const CudaSortedSet &o) const { return !(*this == o); }

    // host iteration over a copy (ascending)
    const SetElement *begin() const { fetch(); return host_.data(); }
    const SetElement *end() const { fetch(); return host_.data() + host_.size(); }

    // batched forms: this set against many sets, or against the neighbourhoods of the members of `members`
    std::vector<uint64_t> intersect_count_many(const std::vector<const CudaSortedSet *> &others) const {
        std::vector<gmsb_set_t> hs;
        for (auto *o : others) hs.push_back(o->h_);
        std::vector<uint64_t> out(others.size());
        check(gmsb_set_op_count_many(GMSB_SET_INTERSECT, h_, (int64_t)hs.size(), hs.data(), out.data()));
        return out;
    }
    std::vector<uint64_t> intersect_count_neighbourhoods(const CudaSetGraph &g, const CudaSortedSet &members) const {
        std::vector<uint64_t> out(members.cardinality());
        check(gmsb_set_op_count_neighbourhoods(GMSB_SET_INTERSECT, h_, g.handle(), members.h_, out.data()));
        return out;
    }
    gmsb_set_t handle() const { return h_; }

private:
    CudaSortedSet op(int kind, const CudaSortedSet &o) const {
        gmsb_set_t h = nullptr;
        check(gmsb_set_op(kind, h_, o.h_, &h));
        return CudaSortedSet(h);
    }
    void op_inplace(int kind, const CudaSortedSet &o) { check(gmsb_set_op_inplace(kind, h_, o.h_)); fetched_ = false; }
    size_t op_count(int kind, const CudaSortedSet &o) const {
        uint64_t c = 0;
        check(gmsb_set_op_count(kind, h_, o.h_, &c));
        return (size_t)c;
    }
    void fetch() const {
        if (fetched_) return;
        host_.assign(cardinality(), 0);
        if (!host_.empty()) check(gmsb_set_to_host(h_, host_.data()));
        fetched_ = true;
    }
    void release() { if (h_) gmsb_set_free(h_); h_ = nullptr; }

    gmsb_set_t h_ = nullptr;
    mutable std::vector<SetElement> host_;
    mutable bool fetched_ = false;
};

}  // namespace gms_b200
