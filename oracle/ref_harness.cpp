// TEST INFRASTRUCTURE ONLY -- not product code, never linked into libgeneakit_b200.so.
//
// C-ABI harness around the UNMODIFIED reference C++ (compiled in place from
// $(GK_REFERENCE)/geneakit/lib/{create,identify,extract,compute}.cpp by oracle/Makefile
// into oracle/_ref/libgkref.so).  It replaces the nanobind module `cgeneakit`
// (geneakit/lib/cgeneakit.cpp), which cannot be built here because nanobind is absent.
// Only tests/, bench.py's reference / cpu_baseline legs and the golden-vector generator
// (oracle/make_golden.py) load this library.
//
// Each entry point is a thin wrapper over one reference function:
//   gkref_load_vectors  -> load_pedigree_from_vectors   (create.cpp:211-218)
//   gkref_load_file     -> load_pedigree_from_file      (create.cpp:202-208)
//   gkref_flatten       -> Pedigree::ids / Individual::rank (pedigree.hpp:34-44)
//   gkref_proband_ids   -> get_proband_ids              (identify.cpp:55-65)
//   gkref_founder_ids   -> get_founder_ids              (identify.cpp:28-38)
//   gkref_phi           -> compute_kinships             (compute.cpp:638-700)
//   gkref_gc            -> compute_genetic_contributions(compute.cpp:832-862)
//   gkref_required_gb   -> get_required_memory_for_kinships (compute.cpp:591-635)
//   gkref_f             -> compute_inbreedings          (compute.cpp:725-814)
//   gkref_sparse        -> compute_kinships_sparse      (compute.cpp:272-487)
//   gkref_cut_sizes     -> cut_vertices                 (compute.cpp:119-135)
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "create.hpp"
#include "compute.hpp"

static thread_local std::string g_err;

extern "C" {

const char *gkref_last_error(void) { return g_err.c_str(); }

void *gkref_load_vectors(int n, const int *ids, const int *fathers,
                         const int *mothers, const int *sexes, int sorted) {
    try {
        std::vector<int> a(ids, ids + n), b(fathers, fathers + n),
            c(mothers, mothers + n), d(sexes, sexes + n);
        Pedigree<> *p = new Pedigree<>();
        *p = load_pedigree_from_vectors(a, b, c, d, sorted != 0);
        return p;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void *gkref_load_file(const char *path, int sorted) {
    try {
        Pedigree<> *p = new Pedigree<>();
        *p = load_pedigree_from_file(std::string(path), sorted != 0);
        return p;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void gkref_free(void *h) { delete static_cast<Pedigree<> *>(h); }

int gkref_size(void *h) { return (int) static_cast<Pedigree<> *>(h)->ids.size(); }

// Rank-ordered flat view: ids[k], father/mother RANK (-1 unknown), sex.
void gkref_flatten(void *h, int *ids, int *father_rank, int *mother_rank, int *sex) {
    Pedigree<> &p = *static_cast<Pedigree<> *>(h);
    int k = 0;
    for (const int id : p.ids) {
        const Individual<> *ind = p.individuals.at(id);
        ids[k] = id;
        father_rank[k] = ind->father ? ind->father->rank : -1;
        mother_rank[k] = ind->mother ? ind->mother->rank : -1;
        sex[k] = (int) ind->sex;
        ++k;
    }
}

int gkref_proband_ids(void *h, int *out) {
    std::vector<int> v = get_proband_ids(*static_cast<Pedigree<> *>(h));
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(int));
    return (int) v.size();
}

int gkref_founder_ids(void *h, int *out) {
    std::vector<int> v = get_founder_ids(*static_cast<Pedigree<> *>(h));
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(int));
    return (int) v.size();
}

// out: m*m doubles, row-major. m == 0 -> default probands (then out must hold |pro|^2).
int gkref_phi(void *h, int m, const int *pro, double *out, int verbose) {
    try {
        std::vector<int> ids(pro, pro + m);
        Matrix<double> k = compute_kinships(*static_cast<Pedigree<> *>(h), ids, verbose != 0);
        std::memcpy(out, k.data(), k.rows() * k.cols() * sizeof(double));
        return 0;
    } catch (const std::out_of_range &e) {
        g_err = e.what();
        return 1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 2;
    }
}

int gkref_gc(void *h, int m, const int *pro, int a, const int *anc, double *out) {
    try {
        std::vector<int> p(pro, pro + m), q(anc, anc + a);
        Matrix<double> k = compute_genetic_contributions(*static_cast<Pedigree<> *>(h), p, q);
        std::memcpy(out, k.data(), k.rows() * k.cols() * sizeof(double));
        return 0;
    } catch (const std::out_of_range &e) {
        g_err = e.what();
        return 1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 2;
    }
}

// gen.f: compute_inbreedings (compute.cpp:725-814).  out: m doubles.
int gkref_f(void *h, int m, const int *pro, double *out) {
    try {
        std::vector<int> ids(pro, pro + m);
        std::vector<double> f = compute_inbreedings(*static_cast<Pedigree<> *>(h), ids);
        std::memcpy(out, f.data(), f.size() * sizeof(double));
        return 0;
    } catch (const std::out_of_range &e) {
        g_err = e.what();
        return 1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 2;
    }
}

// gen.phi(sparse=True): compute_kinships_sparse (compute.cpp:272-487), CSR branch.  First call with data == NULL returns
// nnz; the second fills data (nnz floats), indices (nnz ints), indptr (m + 1 int64).
long long gkref_sparse(void *h, int m, const int *pro, float *data, int *indices, long long *indptr) {
    try {
        std::vector<int> ids(pro, pro + m);
        SparseResult r = compute_kinships_sparse(*static_cast<Pedigree<> *>(h), ids, false, false);
        if (data) {
            std::memcpy(data, r.data.data(), r.data.size() * sizeof(float));
            std::memcpy(indices, r.indices.data(), r.indices.size() * sizeof(int));
            for (size_t i = 0; i < r.indptr.size(); i++) indptr[i] = (long long) r.indptr[i];
        }
        return (long long) r.data.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

double gkref_required_gb(void *h, int m, const int *pro) {
    try {
        std::vector<int> ids(pro, pro + m);
        return get_required_memory_for_kinships(*static_cast<Pedigree<> *>(h), ids);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1.0;
    }
}

// Sizes of the vertex cuts the reference builds for these probands; returns #cuts.
int gkref_cut_sizes(void *h, int m, const int *pro, int *sizes, int cap) {
    try {
        Pedigree<> &ped = *static_cast<Pedigree<> *>(h);
        std::vector<int> ids(pro, pro + m);
        if (ids.empty()) ids = get_proband_ids(ped);
        Pedigree<Indices> kp(ped);
        std::vector<Individual<Indices> *> probands;
        for (const int id : ids) probands.push_back(kp.individuals.at(id));
        std::vector<std::vector<Individual<Indices> *>> cuts = cut_vertices(probands);
        for (int i = 0; i < (int) cuts.size() && i < cap; i++) sizes[i] = (int) cuts[i].size();
        return (int) cuts.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

void gkref_set_threads(int t) { omp_set_num_threads(t); }
int gkref_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
