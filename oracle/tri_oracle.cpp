// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product
// path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library, and only as the checker / CPU baseline.
//
// PARITY UNPINNED for this file: the reference (dewan1988/FLUID-LLM) holds no golden vectors
// for this boundary and its arithmetic lives in a third-party module that is not vendored in
// the reference tree and not installed here: matplotlib 3.8.2 (environemnt.yml:166), C++
// extension `matplotlib._tri`.  What follows is a fresh restatement of that module's
// *published algorithm* as the reference calls it:
//
//   reference call site                          matplotlib._tri entry restated here
//   src/dataloader/mesh_utils.py:103             Triangulation(x, y, triangles)   -> fo_tri_create
//   src/dataloader/mesh_utils.py:104             get_trifinder()(grid_x, grid_y)  -> fo_trifinder_init
//                                                 TrapezoidMapTriFinder.find_many -> fo_trifinder_find_many
//   src/_triinterpolate.py:262-263               calculate_plane_coefficients(z)  -> fo_tri_plane_coefficients
//
// Semantics restated:
//   * correct_triangles: a triangle whose signed area (p1-p0)x(p2-p0) is negative gets its
//     vertices 1 and 2 swapped, so all later maths sees counter-clockwise triangles.
//   * calculate_neighbors: neighbour of (tri, edge k) is the triangle holding the reversed
//     directed edge (v[k+1] -> v[k]), or -1.
//   * TrapezoidMapTriFinder: randomised incremental trapezoidal map (de Berg et al., ch. 6)
//     over every mesh edge, lexicographic (x, then y) point order, edges shuffled with
//     std::mt19937(1234); queries walk the DAG and stop early on an exact vertex hit
//     (-> lowest-index triangle that lists the vertex) or an exact on-edge hit
//     (-> triangle above the edge, else the one below).
//   * plane coefficients: z = a x + b y + c through the three (corrected-order) vertices,
//     fp64, with the colinear pseudo-inverse branch.
//
// A second, independent locator (fo_rule_find_many) evaluates the *stated tie-break rule*
// directly (brute force over triangles, optionally bucketed); the tests assert that the
// trapezoid map and the rule agree on every test mesh.  The CUDA product implements the rule.
//
// Build: g++ -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: matplotlib's x86-64
// wheels are baseline SSE2 builds, so every product/sum below is rounded separately).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <unordered_map>
#include <vector>

namespace {

struct Pt {
    double x, y;
    int tri;  // lowest-index triangle that lists this point, -1 if none
};

// a is lexicographically right of b (x first, then y)
inline bool right_of(const Pt& a, const Pt& b) { return a.x == b.x ? a.y > b.y : a.x > b.x; }
inline bool right_of_xy(double ax, double ay, const Pt& b) { return ax == b.x ? ay > b.y : ax > b.x; }

struct Edge {
    const Pt* left;
    const Pt* right;
    int tri_below, tri_above;
    const Pt* pt_below;  // third vertex of the triangle below (or null)
    const Pt* pt_above;  // third vertex of the triangle above (or null)

    // sign of (p - left) x (right - left); >0 : p below the edge, <0 : above, 0 : on its line
    int orient(double px, double py) const {
        double a = (px - left->x) * (right->y - left->y);
        double b = (py - left->y) * (right->x - left->x);
        double cz = a - b;
        return cz > 0.0 ? +1 : (cz < 0.0 ? -1 : 0);
    }
    double slope() const { return (right->y - left->y) / (right->x - left->x); }
    bool has_point(const Pt* p) const { return p == left || p == right; }
};

struct Node;

struct Trap {
    const Pt* left;
    const Pt* right;
    const Edge* below;
    const Edge* above;
    Trap *ll = nullptr, *lr = nullptr, *ul = nullptr, *ur = nullptr;
    Node* node = nullptr;
    Trap(const Pt* l, const Pt* r, const Edge* b, const Edge* a) : left(l), right(r), below(b), above(a) {}
    void set_ll(Trap* t) { ll = t; if (t) t->lr = this; }
    void set_lr(Trap* t) { lr = t; if (t) t->ll = this; }
    void set_ul(Trap* t) { ul = t; if (t) t->ur = this; }
    void set_ur(Trap* t) { ur = t; if (t) t->ul = this; }
};

enum NodeType { XNODE, YNODE, LEAF };

struct Node {
    NodeType type;
    const Pt* point = nullptr;    // XNODE
    const Edge* edge = nullptr;   // YNODE
    Trap* trap = nullptr;         // LEAF
    Node* a = nullptr;            // XNODE: left   YNODE: below
    Node* b = nullptr;            // XNODE: right  YNODE: above
    std::vector<Node*> parents;
};

struct Triangulation {
    int npoints = 0, ntri = 0;
    std::vector<double> x, y;
    std::vector<int> tris;       // corrected orientation, (ntri,3)
    std::vector<int> neighbors;  // (ntri,3)
    // trapezoid map
    std::vector<Pt> points;
    std::vector<Edge> edges;
    std::vector<std::unique_ptr<Node>> node_arena;
    std::vector<std::unique_ptr<Trap>> trap_arena;
    Node* root = nullptr;
    bool finder_ready = false;
    // bucket grid for the rule-based locator
    bool buckets_ready = false;
    int bnx = 0, bny = 0;
    double bx0 = 0, by0 = 0, bdx = 1, bdy = 1;
    std::vector<int> bstart, bitems;

    Node* new_node(NodeType t) {
        node_arena.emplace_back(new Node());
        node_arena.back()->type = t;
        return node_arena.back().get();
    }
    Trap* new_trap(const Pt* l, const Pt* r, const Edge* b, const Edge* a) {
        trap_arena.emplace_back(new Trap(l, r, b, a));
        return trap_arena.back().get();
    }
    Node* leaf_for(Trap* t) {
        if (!t->node) {
            Node* n = new_node(LEAF);
            n->trap = t;
            t->node = n;
        }
        return t->node;
    }
    Node* make_x(const Pt* p, Node* l, Node* r) {
        Node* n = new_node(XNODE);
        n->point = p; n->a = l; n->b = r;
        l->parents.push_back(n); r->parents.push_back(n);
        return n;
    }
    Node* make_y(const Edge* e, Node* below, Node* above) {
        Node* n = new_node(YNODE);
        n->edge = e; n->a = below; n->b = above;
        below->parents.push_back(n); above->parents.push_back(n);
        return n;
    }
    void replace(Node* old_node, Node* repl) {
        if (old_node == root) { root = repl; return; }
        for (Node* p : old_node->parents) {
            if (p->a == old_node) p->a = repl;
            if (p->b == old_node) p->b = repl;
            repl->parents.push_back(p);
        }
        old_node->parents.clear();
    }
};

// ---- matplotlib Triangulation::correct_triangles + calculate_neighbors ----------------------
void correct_and_neighbors(Triangulation& T) {
    for (int t = 0; t < T.ntri; ++t) {
        int* v = &T.tris[3 * t];
        double dx1 = T.x[v[1]] - T.x[v[0]], dy1 = T.y[v[1]] - T.y[v[0]];
        double dx2 = T.x[v[2]] - T.x[v[0]], dy2 = T.y[v[2]] - T.y[v[0]];
        double cz = dx1 * dy2 - dy1 * dx2;
        if (cz < 0.0) std::swap(v[1], v[2]);
    }
    T.neighbors.assign(3 * (size_t)T.ntri, -1);
    std::unordered_map<uint64_t, int> open;  // directed edge (start,end) -> tri*3+k waiting for its twin
    open.reserve((size_t)T.ntri * 3);
    for (int t = 0; t < T.ntri; ++t) {
        for (int k = 0; k < 3; ++k) {
            uint32_t s = (uint32_t)T.tris[3 * t + k], e = (uint32_t)T.tris[3 * t + (k + 1) % 3];
            auto it = open.find(((uint64_t)e << 32) | s);
            if (it == open.end()) {
                open[((uint64_t)s << 32) | e] = 3 * t + k;
            } else {
                T.neighbors[3 * t + k] = it->second / 3;
                T.neighbors[it->second] = t;
                open.erase(it);
            }
        }
    }
}

// ---- trapezoid map -------------------------------------------------------------------------
// Walk used while inserting an edge: find the leaf that holds the start of `e`.
Node* search_edge(Node* n, const Edge& e) {
    for (;;) {
        if (n->type == LEAF) return n;
        if (n->type == XNODE) {
            if (e.left == n->point) n = n->b;
            else n = right_of(*e.left, *n->point) ? n->b : n->a;
        } else {
            const Edge* y = n->edge;
            if (e.left == y->left) {           // shared left end point: compare slopes
                double se = e.slope(), sy = y->slope();
                if (se == sy) {
                    if (y->tri_above == e.tri_below) n = n->b;
                    else if (y->tri_below == e.tri_above) n = n->a;
                    else return nullptr;
                } else n = (se > sy) ? n->b : n->a;
            } else if (e.right == y->right) {  // shared right end point
                double se = e.slope(), sy = y->slope();
                if (se == sy) {
                    if (y->tri_above == e.tri_below) n = n->b;
                    else if (y->tri_below == e.tri_above) n = n->a;
                    else return nullptr;
                } else n = (se > sy) ? n->a : n->b;
            } else {
                int o = y->orient(e.left->x, e.left->y);
                if (o == 0) {                  // start of e lies on y: use the opposite vertices
                    if (y->pt_above && e.has_point(y->pt_above)) o = -1;
                    else if (y->pt_below && e.has_point(y->pt_below)) o = +1;
                    else return nullptr;
                }
                n = (o < 0) ? n->b : n->a;
            }
        }
    }
}

bool crossed_traps(Triangulation& T, const Edge& e, std::vector<Trap*>& out) {
    out.clear();
    Node* n = search_edge(T.root, e);
    if (!n) return false;
    Trap* t = n->trap;
    out.push_back(t);
    while (right_of(*e.right, *t->right)) {
        int o = e.orient(t->right->x, t->right->y);
        if (o == 0) {
            if (e.pt_below == t->right) o = +1;
            else if (e.pt_above == t->right) o = -1;
            else return false;
        }
        t = (o == -1) ? t->lr : t->ur;  // right end of t above e -> continue through its lower-right
        if (!t) return false;
        out.push_back(t);
    }
    return true;
}

bool insert_edge(Triangulation& T, const Edge& e) {
    std::vector<Trap*> row;
    if (!crossed_traps(T, e, row)) return false;
    const Pt* p = e.left;
    const Pt* q = e.right;
    Trap *prev_old = nullptr, *prev_below = nullptr, *prev_above = nullptr;
    size_t n = row.size();
    for (size_t i = 0; i < n; ++i) {
        Trap* old = row[i];
        bool first = (i == 0), last = (i == n - 1);
        bool cut_left = first && p != old->left;
        bool cut_right = last && q != old->right;
        Trap *lt = nullptr, *rt = nullptr, *below = nullptr, *above = nullptr;
        if (first && last) {
            if (cut_left) lt = T.new_trap(old->left, p, old->below, old->above);
            below = T.new_trap(p, q, old->below, &e);
            above = T.new_trap(p, q, &e, old->above);
            if (cut_right) rt = T.new_trap(q, old->right, old->below, old->above);
            if (cut_left) {
                lt->set_ll(old->ll); lt->set_ul(old->ul);
                lt->set_lr(below); lt->set_ur(above);
            } else {
                below->set_ll(old->ll); above->set_ul(old->ul);
            }
            if (cut_right) {
                rt->set_lr(old->lr); rt->set_ur(old->ur);
                below->set_lr(rt); above->set_ur(rt);
            } else {
                below->set_lr(old->lr); above->set_ur(old->ur);
            }
        } else if (first) {
            if (cut_left) lt = T.new_trap(old->left, p, old->below, old->above);
            below = T.new_trap(p, old->right, old->below, &e);
            above = T.new_trap(p, old->right, &e, old->above);
            if (cut_left) {
                lt->set_ll(old->ll); lt->set_ul(old->ul);
                lt->set_lr(below); lt->set_ur(above);
            } else {
                below->set_ll(old->ll); above->set_ul(old->ul);
            }
            below->set_lr(old->lr); above->set_ur(old->ur);
        } else {
            const Pt* rgt = last ? q : old->right;
            if (prev_below->below == old->below) { below = prev_below; below->right = rgt; }
            else below = T.new_trap(old->left, rgt, old->below, &e);
            if (prev_above->above == old->above) { above = prev_above; above->right = rgt; }
            else above = T.new_trap(old->left, rgt, &e, old->above);
            if (last && cut_right) {
                rt = T.new_trap(q, old->right, old->below, old->above);
                rt->set_lr(old->lr); rt->set_ur(old->ur);
            }
            if (below != prev_below) {
                below->set_ul(prev_below);
                below->set_ll(old->ll == prev_old ? prev_below : old->ll);
            }
            if (above != prev_above) {
                above->set_ll(prev_above);
                above->set_ul(old->ul == prev_old ? prev_above : old->ul);
            }
            if (last && cut_right) { below->set_lr(rt); above->set_ur(rt); }
            else { below->set_lr(old->lr); above->set_ur(old->ur); }
        }
        Node* top = T.make_y(&e, T.leaf_for(below), T.leaf_for(above));
        if (rt) top = T.make_x(q, top, T.leaf_for(rt));
        if (lt) top = T.make_x(p, T.leaf_for(lt), top);
        T.replace(old->node, top);
        if (!last) { prev_old = old; prev_below = below; prev_above = above; }
    }
    return true;
}

int build_finder(Triangulation& T) {
    T.node_arena.clear(); T.trap_arena.clear(); T.edges.clear(); T.points.clear();
    T.root = nullptr; T.finder_ready = false;
    int np = T.npoints;
    T.points.resize(np + 4);
    double lox = 0, loy = 0, hix = 0, hiy = 0;
    for (int i = 0; i < np; ++i) {
        double px = T.x[i], py = T.y[i];
        if (px == -0.0) px = 0.0;  // normalise signed zero
        if (py == -0.0) py = 0.0;
        T.points[i] = Pt{px, py, -1};
        if (i == 0) { lox = hix = px; loy = hiy = py; }
        else { lox = std::min(lox, px); hix = std::max(hix, px); loy = std::min(loy, py); hiy = std::max(hiy, py); }
    }
    if (np == 0) { lox = loy = 0.0; hix = hiy = 1.0; }
    else {
        double ex = (hix - lox) * 0.1, ey = (hiy - loy) * 0.1;
        lox -= ex; hix += ex; loy -= ey; hiy += ey;
    }
    T.points[np] = Pt{lox, loy, -1};      // SW
    T.points[np + 1] = Pt{hix, loy, -1};  // SE
    T.points[np + 2] = Pt{lox, hiy, -1};  // NW
    T.points[np + 3] = Pt{hix, hiy, -1};  // NE
    const Pt* P = T.points.data();
    T.edges.reserve(2 + 3 * (size_t)T.ntri);
    T.edges.push_back(Edge{P + np, P + np + 1, -1, -1, nullptr, nullptr});
    T.edges.push_back(Edge{P + np + 2, P + np + 3, -1, -1, nullptr, nullptr});
    for (int t = 0; t < T.ntri; ++t) {
        for (int k = 0; k < 3; ++k) {
            Pt* start = &T.points[T.tris[3 * t + k]];
            Pt* end = &T.points[T.tris[3 * t + (k + 1) % 3]];
            Pt* other = &T.points[T.tris[3 * t + (k + 2) % 3]];
            int nb = T.neighbors[3 * t + k];
            if (right_of(*end, *start)) {
                const Pt* below_pt = nullptr;
                if (nb != -1) {
                    // vertex of the neighbour that is not on the shared edge
                    for (int kk = 0; kk < 3; ++kk) {
                        int v = T.tris[3 * nb + kk];
                        if (v != T.tris[3 * t + k] && v != T.tris[3 * t + (k + 1) % 3]) below_pt = P + v;
                    }
                }
                T.edges.push_back(Edge{start, end, nb, t, below_pt, other});
            } else if (nb == -1) {
                T.edges.push_back(Edge{end, start, t, -1, other, nullptr});
            }
            if (start->tri == -1) start->tri = t;
        }
    }
    T.root = T.leaf_for(T.new_trap(P + np, P + np + 1, &T.edges[0], &T.edges[1]));
    std::mt19937 rng(1234);
    std::shuffle(T.edges.begin() + 2, T.edges.end(), rng);
    for (size_t i = 2; i < T.edges.size(); ++i)
        if (!insert_edge(T, T.edges[i])) return 1;
    T.finder_ready = true;
    return 0;
}

int find_one(const Triangulation& T, double qx, double qy) {
    const Node* n = T.root;
    for (;;) {
        if (n->type == LEAF) return n->trap->below->tri_above;
        if (n->type == XNODE) {
            const Pt& p = *n->point;
            if (qx == p.x && qy == p.y) return p.tri;
            n = right_of_xy(qx, qy, p) ? n->b : n->a;
        } else {
            int o = n->edge->orient(qx, qy);
            if (o == 0) return n->edge->tri_above != -1 ? n->edge->tri_above : n->edge->tri_below;
            n = (o < 0) ? n->b : n->a;
        }
    }
}

// ---- the stated tie-break rule, evaluated directly ----------------------------------------
// Acceptance of query (qx,qy) by triangle t (corrected order):
//   (i)  q equals one of t's vertices, or
//   (ii) for each edge k (start=v[k], end=v[k+1]) with lexicographically ordered end points
//        (l, r):  s = (q-l) x (r-l);  if end is right of start (t lies above the edge) need
//        s <= 0;  otherwise (t lies below) need s > 0, or s >= 0 when the edge has no neighbour.
// Result: the lowest-index accepting triangle, else -1.
inline bool rule_accepts(const Triangulation& T, int t, double qx, double qy) {
    const int* v = &T.tris[3 * t];
    double vx[3], vy[3];
    for (int k = 0; k < 3; ++k) {
        vx[k] = T.x[v[k]]; vy[k] = T.y[v[k]];
        if (vx[k] == -0.0) vx[k] = 0.0;
        if (vy[k] == -0.0) vy[k] = 0.0;
        if (qx == vx[k] && qy == vy[k]) return true;
    }
    for (int k = 0; k < 3; ++k) {
        int k1 = (k + 1) % 3;
        bool end_right = (vx[k1] == vx[k]) ? (vy[k1] > vy[k]) : (vx[k1] > vx[k]);
        double lx, ly, rx, ry;
        if (end_right) { lx = vx[k]; ly = vy[k]; rx = vx[k1]; ry = vy[k1]; }
        else { lx = vx[k1]; ly = vy[k1]; rx = vx[k]; ry = vy[k]; }
        double a = (qx - lx) * (ry - ly);
        double b = (qy - ly) * (rx - lx);
        double s = a - b;
        if (end_right) { if (!(s <= 0.0)) return false; }
        else if (T.neighbors[3 * t + k] == -1) { if (!(s >= 0.0)) return false; }
        else { if (!(s > 0.0)) return false; }
    }
    return true;
}

void build_buckets(Triangulation& T) {
    double lox = 0, loy = 0, hix = 1, hiy = 1;
    for (int i = 0; i < T.npoints; ++i) {
        if (i == 0) { lox = hix = T.x[i]; loy = hiy = T.y[i]; }
        else { lox = std::min(lox, T.x[i]); hix = std::max(hix, T.x[i]); loy = std::min(loy, T.y[i]); hiy = std::max(hiy, T.y[i]); }
    }
    double w = std::max(hix - lox, 1e-300), h = std::max(hiy - loy, 1e-300);
    double target = std::sqrt((double)std::max(T.ntri, 1) / 2.0);
    double aspect = w / h;
    T.bnx = std::max(1, std::min(4096, (int)std::ceil(target * std::sqrt(aspect))));
    T.bny = std::max(1, std::min(4096, (int)std::ceil(target / std::sqrt(aspect))));
    T.bx0 = lox; T.by0 = loy; T.bdx = w / T.bnx; T.bdy = h / T.bny;
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    std::vector<int> count((size_t)T.bnx * T.bny + 1, 0);
    auto range = [&](int t, int& i0, int& i1, int& j0, int& j1) {
        const int* v = &T.tris[3 * t];
        double x0 = std::min({T.x[v[0]], T.x[v[1]], T.x[v[2]]}), x1 = std::max({T.x[v[0]], T.x[v[1]], T.x[v[2]]});
        double y0 = std::min({T.y[v[0]], T.y[v[1]], T.y[v[2]]}), y1 = std::max({T.y[v[0]], T.y[v[1]], T.y[v[2]]});
        i0 = clampi((int)std::floor((x0 - T.bx0) / T.bdx) - 1, 0, T.bnx - 1);
        i1 = clampi((int)std::floor((x1 - T.bx0) / T.bdx) + 1, 0, T.bnx - 1);
        j0 = clampi((int)std::floor((y0 - T.by0) / T.bdy) - 1, 0, T.bny - 1);
        j1 = clampi((int)std::floor((y1 - T.by0) / T.bdy) + 1, 0, T.bny - 1);
    };
    for (int t = 0; t < T.ntri; ++t) {
        int i0, i1, j0, j1; range(t, i0, i1, j0, j1);
        for (int i = i0; i <= i1; ++i) for (int j = j0; j <= j1; ++j) count[(size_t)i * T.bny + j + 1]++;
    }
    for (size_t i = 1; i < count.size(); ++i) count[i] += count[i - 1];
    T.bstart = count;
    T.bitems.assign(count.back(), 0);
    std::vector<int> cur(count.begin(), count.end() - 1);
    for (int t = 0; t < T.ntri; ++t) {  // ascending t => every bucket list is ascending
        int i0, i1, j0, j1; range(t, i0, i1, j0, j1);
        for (int i = i0; i <= i1; ++i) for (int j = j0; j <= j1; ++j) T.bitems[cur[(size_t)i * T.bny + j]++] = t;
    }
    T.buckets_ready = true;
}

}  // namespace

extern "C" {

// x, y: float64[npoints]; triangles: int32[ntri*3].  Copies everything; applies
// correct_triangles and calculate_neighbors (matplotlib does so lazily on first C++ use).
void* fo_tri_create(const double* x, const double* y, int npoints, const int* triangles, int ntri) {
    for (int i = 0; i < 3 * ntri; ++i)
        if (triangles[i] < 0 || triangles[i] >= npoints) return nullptr;
    Triangulation* T = new Triangulation();
    T->npoints = npoints; T->ntri = ntri;
    T->x.assign(x, x + npoints); T->y.assign(y, y + npoints);
    T->tris.assign(triangles, triangles + 3 * (size_t)ntri);
    correct_and_neighbors(*T);
    return T;
}
void fo_tri_destroy(void* h) { delete (Triangulation*)h; }
void fo_tri_get_triangles(void* h, int* out) {
    Triangulation* T = (Triangulation*)h;
    std::memcpy(out, T->tris.data(), sizeof(int) * T->tris.size());
}
void fo_tri_get_neighbors(void* h, int* out) {
    Triangulation* T = (Triangulation*)h;
    std::memcpy(out, T->neighbors.data(), sizeof(int) * T->neighbors.size());
}

// matplotlib Triangulation::calculate_plane_coefficients (fp64; out: float64[ntri*3])
void fo_tri_plane_coefficients(void* h, const double* z, double* out) {
    Triangulation* T = (Triangulation*)h;
    for (int t = 0; t < T->ntri; ++t) {
        const int* v = &T->tris[3 * t];
        double p0x = T->x[v[0]], p0y = T->y[v[0]], p0z = z[v[0]];
        double ax = T->x[v[1]] - p0x, ay = T->y[v[1]] - p0y, az = z[v[1]] - p0z;  // side01
        double bx = T->x[v[2]] - p0x, by = T->y[v[2]] - p0y, bz = z[v[2]] - p0z;  // side02
        double nx = ay * bz - az * by;
        double ny = az * bx - ax * bz;
        double nz = ax * by - ay * bx;
        if (nz == 0.0) {
            double sum2 = ax * ax + ay * ay + bx * bx + by * by;
            double a = (ax * az + bx * bz) / sum2;
            double b = (ay * az + by * bz) / sum2;
            out[3 * t] = a; out[3 * t + 1] = b; out[3 * t + 2] = p0z - a * p0x - b * p0y;
        } else {
            out[3 * t] = -nx / nz;
            out[3 * t + 1] = -ny / nz;
            out[3 * t + 2] = (nx * p0x + ny * p0y + nz * p0z) / nz;
        }
    }
}

// 0 = ok, 1 = "Triangulation is invalid"
int fo_trifinder_init(void* h) { return build_finder(*(Triangulation*)h); }

void fo_trifinder_find_many(void* h, const double* x, const double* y, long n, int* out) {
    Triangulation* T = (Triangulation*)h;
    for (long i = 0; i < n; ++i) out[i] = find_one(*T, x[i], y[i]);
}

// node count and maximum depth of the search DAG (diagnostics only)
void fo_trifinder_stats(void* h, long* n_nodes, long* n_traps) {
    Triangulation* T = (Triangulation*)h;
    *n_nodes = (long)T->node_arena.size();
    *n_traps = (long)T->trap_arena.size();
}

// rule-based locator. bucketed = 0: test every triangle (O(n*ntri)); 1: uniform bucket grid.
void fo_rule_find_many(void* h, const double* x, const double* y, long n, int bucketed, int* out) {
    Triangulation* T = (Triangulation*)h;
    if (bucketed && !T->buckets_ready) build_buckets(*T);
    for (long i = 0; i < n; ++i) {
        int res = -1;
        if (!bucketed) {
            for (int t = 0; t < T->ntri; ++t)
                if (rule_accepts(*T, t, x[i], y[i])) { res = t; break; }
        } else {
            int bi = (int)std::floor((x[i] - T->bx0) / T->bdx), bj = (int)std::floor((y[i] - T->by0) / T->bdy);
            if (bi >= -1 && bi <= T->bnx && bj >= -1 && bj <= T->bny) {
                bi = std::min(std::max(bi, 0), T->bnx - 1); bj = std::min(std::max(bj, 0), T->bny - 1);
                size_t b = (size_t)bi * T->bny + bj;
                for (int k = T->bstart[b]; k < T->bstart[b + 1]; ++k)
                    if (rule_accepts(*T, T->bitems[k], x[i], y[i])) { res = T->bitems[k]; break; }
            }
        }
        out[i] = res;
    }
}

// Whole-frame CPU baseline helper (src/dataloader/mesh_utils.py:82-91 restated in C for timing):
// per channel, plane coefficients for all triangles then the gather z = a x + b y + c over the
// grid cells with tri != -1; masked cells -> 0.  gx, gy: float32 grid coordinates (n cells).
void fo_to_grid(void* h, const float* val, const float* gx, const float* gy, const int* tri_index, long n,
                float* data, unsigned char* mask, double* plane_scratch, double* z_scratch) {
    Triangulation* T = (Triangulation*)h;
    for (int i = 0; i < T->npoints; ++i) z_scratch[i] = (double)val[i];
    fo_tri_plane_coefficients(h, z_scratch, plane_scratch);
    for (long i = 0; i < n; ++i) {
        int t = tri_index[i];
        if (t < 0) { data[i] = 0.0f; mask[i] = 1; continue; }
        double ax = plane_scratch[3 * t] * (double)gx[i];
        double by = plane_scratch[3 * t + 1] * (double)gy[i];
        double s = ax + by;
        float r = (float)(s + plane_scratch[3 * t + 2]);
        bool bad = !std::isfinite(r);
        data[i] = bad ? 0.0f : r;
        mask[i] = bad ? 1 : 0;
    }
}

}  // extern "C"
