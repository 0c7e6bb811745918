// oracle/shim/ref_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry layer over the reference's OWN classes (atom, AtomList, InterAtomList, NewtonMotion, the packers,
// ws::*, configuration::*), compiled in place from /root/reference/src by oracle/Makefile into
// oracle/_ref/libmisa_ref.so. It drives them exactly the way simulation::prepareForStart / simulate /
// collisionStep do (reference src/simulation.cpp:137-145,164-194,208-217), with one thread per MPI rank.
// With -DREF_WITH_CUDA_HOOKS the reference is built in its accelerated configuration (ACCELERATE_ENABLED,
// ARCH_NAME=cuda) and atom::latRho/latDf/latForce call arch_cuda/cuda_hooks.cpp, i.e. the product.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <mpi.h>
#include <comm/thread_world.h>
#include <comm/domain/bcc_domain.h>
#include <eam.h>

#include "atom.h"
#include "newton_motion.h"
#include "system_configuration.h"
#include "world_builder.h"
#include "lattice/ws_utils.h"
#include "utils/mpi_data_types.h"
#include "arch/arch_env.hpp"
#include "arch/hardware_accelerate.hpp"

// ---- the thread "MPI" ---------------------------------------------------------------------------------
namespace shim {
    thread_local int tl_rank = 0;
    thread_local ThreadWorld *tl_world = nullptr;
    static ThreadWorld g_single(1);
}
static shim::ThreadWorld *cur_world() { return shim::tl_world ? shim::tl_world : &shim::g_single; }

extern "C" {
double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int MPI_Abort(MPI_Comm, int code) {
    fprintf(stderr, "MPI_Abort(%d) from the reference code\n", code);
    abort();
}
// sum over the rank threads, in rank order (deterministic)
int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
    shim::ThreadWorld *w = cur_world();
    if (type != MPI_DOUBLE || count > 16) { fprintf(stderr, "shim MPI_Allreduce: unsupported\n"); abort(); }
    const double *s = static_cast<const double *>(send);
    double *r = static_cast<double *>(recv);
    if (w->n == 1) { for (int i = 0; i < count; i++) r[i] = s[i]; return 0; }
    for (int i = 0; i < count; i++) w->reduce[16 * shim::tl_rank + i] = s[i];
    w->barrier.wait();
    for (int i = 0; i < count; i++) {
        double acc = 0.0;
        for (int k = 0; k < w->n; k++) acc += w->reduce[16 * k + i];
        r[i] = acc;
    }
    w->barrier.wait();
    return 0;
}
int MPI_Reduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op op, int, MPI_Comm comm) {
    return MPI_Allreduce(send, recv, count, type, op, comm);
}
int MPI_Initialized(int *flag) { *flag = 0; return 0; }
int MPI_Comm_split_type(MPI_Comm comm, int, int, MPI_Info, MPI_Comm *newcomm) { *newcomm = comm; return 0; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = 0; return 0; }
int MPI_Comm_free(MPI_Comm *) { return 0; }
int MPI_Get_address(const void *location, MPI_Aint *address) { *address = (MPI_Aint)location; return 0; }
int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *newtype) { *newtype = 100; return 0; }
int MPI_Type_commit(MPI_Datatype *) { return 0; }
int MPI_Type_free(MPI_Datatype *) { return 0; }
}

// AtomList::sendlist / recvlist and _atoms are private; AtomDump is a declared friend (atom_list.h:26) that
// lives in the reference's frontend, which is not compiled here.
class AtomDump {
public:
    static AtomElement *atoms(AtomList *l) { return l->_atoms; }
    static std::vector<std::vector<_type_atom_id>> &sendlist(AtomList *l) { return l->sendlist; }
    static std::vector<std::vector<_type_atom_id>> &recvlist(AtomList *l) { return l->recvlist; }
};

size_t ref_dump_rank(AtomList *atom_list, InterAtomList *inter_list, const comm::BccDomain *d, size_t time_step,
                     unsigned char *out, size_t cap); // ref_dump.cpp

struct RefRank {
    comm::BccDomain *dom = nullptr;
    atom *at = nullptr;
    NewtonMotion *nm = nullptr;
};
struct RefWorld {
    int n = 0;
    std::vector<RefRank> ranks;
    pot_eam *pot = nullptr;
    eam *pot_obj = nullptr;
    shim::ThreadWorld *tw = nullptr;
    double comm_time = 0;
};

template<typename Fn>
static void run_all(RefWorld *w, Fn fn) {
    if (w->n == 1) {
        shim::tl_rank = 0;
        shim::tl_world = w->tw;
        fn(0);
        return;
    }
    std::vector<std::thread> th;
    for (int r = 0; r < w->n; r++)
        th.emplace_back([w, r, &fn]() {
            shim::tl_rank = r;
            shim::tl_world = w->tw;
            fn(r);
        });
    for (auto &t : th) t.join();
}

extern "C" {

void *ref_world_create(const long phase_space[3], const int grid_size[3], double lattice_const, double cutoff_radius_factor,
                       const char *setfl_path, double dt) {
    RefWorld *w = new RefWorld();
    w->pot = pot_read_setfl(setfl_path);
    if (!w->pot) { delete w; return nullptr; }
    w->pot_obj = new eam(w->pot);
    w->n = grid_size[0] * grid_size[1] * grid_size[2];
    w->ranks.resize(w->n);
    w->tw = new shim::ThreadWorld(w->n);
    mpi_types::setInterMPIType(); // reference frontend/misa_md.cpp:97
    const int64_t ps[3] = {phase_space[0], phase_space[1], phase_space[2]};
    for (int cx = 0; cx < grid_size[0]; cx++)
        for (int cy = 0; cy < grid_size[1]; cy++)
            for (int cz = 0; cz < grid_size[2]; cz++) {
                const int coord[3] = {cx, cy, cz};
                // simulation::createDomain, reference src/simulation.cpp:41-47
                comm::BccDomain *d = comm::BccDomain::Builder()
                                         .setPhaseSpace(ps)
                                         .setCutoffRadius(cutoff_radius_factor)
                                         .setLatticeConst(lattice_const)
                                         .setGhostSize(static_cast<int>(ceil(cutoff_radius_factor)) + 1)
                                         .localBuild(grid_size, coord);
                RefRank &rk = w->ranks[d->rank];
                rk.dom = d;
                // simulation::createAtoms, reference src/simulation.cpp:62-80
                rk.at = new atom(d);
                rk.at->calcNeighbourIndices(d->cutoff_radius_factor, d->cut_lattice);
                rk.nm = new NewtonMotion(dt);
                AtomElement *a = AtomDump::atoms(rk.at->getAtomList());
                const long n = rk.at->getAtomList()->cap();
                memset(a, 0, sizeof(AtomElement) * n);
                for (long i = 0; i < n; i++) a[i].type = atom_type::INVALID;
            }
#ifdef REF_WITH_CUDA_HOOKS
    if (w->n != 1) { fprintf(stderr, "ref_world_create: the hook build drives one sub-box per process\n"); abort(); }
    archEnvInit();                                        // reference frontend/misa_md.cpp:92
    archAccDomainInit(w->ranks[0].dom);                   // reference src/simulation.cpp:52-54
    archAccNeiOffsetInit(w->ranks[0].at->getNeiOffsets()); // reference src/simulation.cpp:67-69
    archAccPotInit(w->pot_obj);                           // reference src/simulation.cpp:133
#endif
    return w;
}

void ref_world_free(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    if (!w) return;
#ifdef REF_WITH_CUDA_HOOKS
    archEnvFinalize();
#endif
    for (auto &rk : w->ranks) { delete rk.at; delete rk.nm; delete rk.dom; }
    delete w->pot_obj;
    pot_free(w->pot);
    delete w->tw;
    delete w;
}

int ref_accelerated(void) { return isArchAccSupport() ? 1 : 0; }
int ref_n_ranks(void *h) { return static_cast<RefWorld *>(h)->n; }
void *ref_atoms(void *h, int r) { return AtomDump::atoms(static_cast<RefWorld *>(h)->ranks[r].at->getAtomList()); }
long ref_size(void *h, int r) { return static_cast<RefWorld *>(h)->ranks[r].at->getAtomList()->cap(); }
void ref_layout(void *h, int r, int ext[3], int box[3], int ghost[3], int coord[3]) {
    const comm::BccDomain *d = static_cast<RefWorld *>(h)->ranks[r].dom;
    for (int k = 0; k < 3; k++) {
        ext[k] = d->dbx_ghost_extended_lattice_size[k];
        box[k] = d->dbx_sub_box_lattice_size[k];
        ghost[k] = d->dbx_lattice_size_ghost[k];
        coord[k] = d->grid_coord[k];
    }
}

// WorldBuilder::build for a single-sub-box world (its RNG is a process-wide static)
void ref_build_world(void *h, int seed, double t_set, const int ratio[3]) {
    RefWorld *w = static_cast<RefWorld *>(h);
    if (w->n != 1) { fprintf(stderr, "ref_build_world: single sub-box only\n"); abort(); }
    run_all(w, [&](int r) {
        RefRank &rk = w->ranks[r];
        WorldBuilder b;
        b.setDomain(rk.dom).setAtomsContainer(rk.at)
            .setBoxSize(rk.dom->phase_space[0], rk.dom->phase_space[1], rk.dom->phase_space[2])
            .setRandomSeed(seed).setLatticeConst(rk.dom->lattice_const).setTset(t_set).setAlloyRatio(ratio).build();
    });
}

void ref_set_dt(void *h, double dt) {
    RefWorld *w = static_cast<RefWorld *>(h);
    for (auto &rk : w->ranks) rk.nm->setTimestepLength(dt);
}

// simulation::prepareForStart tail, reference src/simulation.cpp:137-145
void ref_prepare(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    run_all(w, [&](int r) {
        RefRank &rk = w->ranks[r];
        double comm = 0;
        rk.at->getAtomList()->exchangeAtomFirst(rk.dom);
        rk.at->clearForce();
        rk.at->computeEam(w->pot_obj, comm);
    });
}

// one iteration of simulation::simulate, reference src/simulation.cpp:164-194
void ref_step(void *h, int n_steps) {
    RefWorld *w = static_cast<RefWorld *>(h);
    run_all(w, [&](int r) {
        RefRank &rk = w->ranks[r];
        for (int s = 0; s < n_steps; s++) {
            double comm = 0;
            rk.nm->firststep(rk.at->getAtomList(), rk.at->getInterList());
            rk.at->decide();
            rk.at->getInterList()->exchangeInter(rk.dom);
            rk.at->getInterList()->borderInter(rk.dom);
            rk.at->getAtomList()->exchangeAtom(rk.dom);
            rk.at->clearForce();
            rk.at->computeEam(w->pot_obj, comm);
            rk.nm->secondstep(rk.at->getAtomList(), rk.at->getInterList());
        }
    });
}

// simulation::collisionStep, reference src/simulation.cpp:208-217
void ref_collision_step(void *h, const int lat[4], const double direction[3], double energy) {
    RefWorld *w = static_cast<RefWorld *>(h);
    run_all(w, [&](int r) {
        RefRank &rk = w->ranks[r];
        double comm = 0;
        rk.at->setv(lat, direction, energy);
        rk.at->getInterList()->exchangeInter(rk.dom);
        rk.at->getInterList()->borderInter(rk.dom);
        rk.at->getAtomList()->exchangeAtom(rk.dom);
        rk.at->clearForce();
        rk.at->computeEam(w->pot_obj, comm);
    });
}

// individual pieces (kernel-level parity and known-answer tests)
void ref_exchange_atom_first(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    run_all(w, [&](int r) { w->ranks[r].at->getAtomList()->exchangeAtomFirst(w->ranks[r].dom); });
}
void ref_clear_force(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    for (auto &rk : w->ranks) rk.at->clearForce();
}
void ref_compute_eam(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    run_all(w, [&](int r) { double comm = 0; w->ranks[r].at->computeEam(w->pot_obj, comm); });
}
void ref_first_step(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    for (auto &rk : w->ranks) rk.nm->firststep(rk.at->getAtomList(), rk.at->getInterList());
}
void ref_second_step(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    for (auto &rk : w->ranks) rk.nm->secondstep(rk.at->getAtomList(), rk.at->getInterList());
}
int ref_decide(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    int f = 0;
    for (auto &rk : w->ranks) f |= rk.at->decide();
    return f;
}
void ref_setv(void *h, const int lat[4], const double direction[3], double energy) {
    RefWorld *w = static_cast<RefWorld *>(h);
    for (auto &rk : w->ranks) rk.at->setv(lat, direction, energy);
}

size_t ref_list_len(void *h, int r, int recv, int index) {
    AtomList *l = static_cast<RefWorld *>(h)->ranks[r].at->getAtomList();
    auto &v = recv ? AtomDump::recvlist(l) : AtomDump::sendlist(l);
    return (size_t)index < v.size() ? v[index].size() : 0;
}
void ref_list_get(void *h, int r, int recv, int index, long *out) {
    AtomList *l = static_cast<RefWorld *>(h)->ranks[r].at->getAtomList();
    auto &v = recv ? AtomDump::recvlist(l) : AtomDump::sendlist(l);
    for (size_t i = 0; i < v[index].size(); i++) out[i] = (long)v[index][i];
}

// NeighbourIndex's protected offset vectors, through a subclass (the reference's own tests do the same,
// tests/unit/neighbour_index_test.cpp:12-26)
class NeiPeek : public NeighbourIndex<AtomElement> {
public:
    static const std::vector<NeiOffset> &get(NeighbourIndex<AtomElement> *n, int which) {
        NeiPeek *p = static_cast<NeiPeek *>(n);
        return which == 0 ? p->nei_even_offsets : which == 1 ? p->nei_odd_offsets : which == 2 ? p->nei_half_even_offsets : p->nei_half_odd_offsets;
    }
};
size_t ref_nei_len(void *h, int r, int which) { return NeiPeek::get(static_cast<RefWorld *>(h)->ranks[r].at->getNeiOffsets(), which).size(); }
void ref_nei_get(void *h, int r, int which, long *out) {
    const auto &v = NeiPeek::get(static_cast<RefWorld *>(h)->ranks[r].at->getNeiOffsets(), which);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
}

size_t ref_n_inter(void *h, int r) { return static_cast<RefWorld *>(h)->ranks[r].at->getInterList()->inter_list.size(); }
size_t ref_n_ghost_inter(void *h, int r) { return static_cast<RefWorld *>(h)->ranks[r].at->getInterList()->inter_ghost_list.size(); }
void ref_get_inter(void *h, int r, void *out) {
    AtomElement *o = static_cast<AtomElement *>(out);
    for (AtomElement &a : static_cast<RefWorld *>(h)->ranks[r].at->getInterList()->inter_list) *o++ = a;
}
void ref_set_inter(void *h, int r, const void *atoms, size_t n) {
    InterAtomList *l = static_cast<RefWorld *>(h)->ranks[r].at->getInterList();
    l->inter_list.clear();
    l->nlocalinter = 0;
    const AtomElement *a = static_cast<const AtomElement *>(atoms);
    for (size_t i = 0; i < n; i++) { AtomElement e = a[i]; l->addInterAtom(e); }
}

// configuration::mvv summed over the sub-boxes; temperature; rescale (reference src/system_configuration.cpp:45-111)
double ref_mvv(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    double e = 0;
    for (auto &rk : w->ranks) e += configuration::mvv(rk.at->getAtomList(), rk.at->getInterList());
    return e;
}
double ref_temperature(void *h) {
    RefWorld *w = static_cast<RefWorld *>(h);
    std::vector<double> t(w->n);
    const comm::BccDomain *d = w->ranks[0].dom;
    const _type_atom_count n_atoms = 2ul * d->phase_space[0] * d->phase_space[1] * d->phase_space[2];
    run_all(w, [&](int r) { t[r] = configuration::temperature(n_atoms, w->ranks[r].at->getAtomList(), w->ranks[r].at->getInterList()); });
    return t[0];
}
void ref_rescale(void *h, double T) {
    RefWorld *w = static_cast<RefWorld *>(h);
    const comm::BccDomain *d = w->ranks[0].dom;
    const _type_atom_count n_atoms = 2ul * d->phase_space[0] * d->phase_space[1] * d->phase_space[2];
    run_all(w, [&](int r) { configuration::rescale(T, n_atoms, w->ranks[r].at->getAtomList(), w->ranks[r].at->getInterList()); });
}

// AtomDump::dump of one sub-box at `time_step` (the reference's own frontend/io sources, see ref_dump.cpp):
// returns the number of bytes of the 72-byte record stream, copied to `out` when it fits
size_t ref_dump(void *h, int r, size_t time_step, void *out, size_t cap) {
    RefRank &rk = static_cast<RefWorld *>(h)->ranks[r];
    return ref_dump_rank(rk.at->getAtomList(), rk.at->getInterList(), rk.dom, time_step, static_cast<unsigned char *>(out), cap);
}

// ws::* (reference src/lattice/ws_utils.cpp) on a free-standing position
unsigned ref_is_out_box(void *h, int r, const double x[3]) {
    AtomElement a;
    memset(&a, 0, sizeof a);
    a.x[0] = x[0]; a.x[1] = x[1]; a.x[2] = x[2];
    return ws::isOutBox(a, static_cast<RefWorld *>(h)->ranks[r].dom);
}
void ref_near_lat_sub_box_coord(void *h, int r, const double x[3], long out[3]) {
    AtomElement a;
    memset(&a, 0, sizeof a);
    a.x[0] = x[0]; a.x[1] = x[1]; a.x[2] = x[2];
    _type_atom_index c[3];
    ws::getNearLatSubBoxCoord(a, static_cast<RefWorld *>(h)->ranks[r].dom, c);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
}

} // extern "C"
