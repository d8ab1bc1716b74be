// Internal declarations shared by the host-side helpers and the CUDA translation units.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace elx {

// One step of the tree walk.  Nodes are visited parent-before-child in "heavy child last" order so
// that the last child can overwrite its parent's slot; at most floor(log2 N)+1 slots are ever live.
struct WalkRec {
  int32_t node;      // preorder index: selects the table, the uniform row and the Philox counter
  int16_t src;       // slot holding the parent's states (slot 0 = root states for the root node)
  int16_t dst;       // slot receiving this node's states, -1 for leaves (written to the output only)
  int32_t leaf_row;  // output row or -1
  int32_t pad;
};
static_assert(sizeof(WalkRec) == 16, "WalkRec must stay 16 bytes");

struct WalkProgram {
  std::vector<WalkRec> recs;
  int n_slots = 0;
};

// Reference draw order is preorder (MarkovProcessAlongTree.hs:88-98); draws are keyed by node id, so
// the visiting order is free.  Returns false (with msg) when the arrays do not describe a preorder tree.
bool compile_walk(int n_nodes, const int32_t *parent, const int32_t *leaf_row, WalkProgram &out, int &n_leaves,
                  std::string &msg);

// One step of the PAIR walk (models with K <= 4 states): the children of a node are drawn in groups of two from ONE joint
// alias row indexed by the parent's state (16 outcomes = state A | state B << 2), which halves the draws per cell.
// Member A is the one whose states stay in registers for the next record (an internal node; the lighter one when both
// are internal), member B is pushed to a slot (internal) or written to its leaf row.  nodeB = -1: a group of one (the
// root, the odd child of a multifurcation, the child of a unary node).
struct PairRec {
  int32_t nodeA, nodeB;
  int16_t src;         // slot holding the parent's states; -1: registers (the previous record's member A, or the root states)
  int8_t dstA, dstB;   // slot receiving A's / B's states, -1: none (A: only nodes with several groups need one)
  int32_t leafA, leafB;  // output rows, -1 for internal nodes (and for an absent B)
};

struct PairProgram {
  std::vector<PairRec> recs;
  int n_slots = 0;
};

// Groups the children of every node (lightest first, the heaviest child in the last group) and orders the records so that
// a record whose parent is the previous record's member A follows it directly.  Expects arrays compile_walk accepted.
void compile_pair_walk(int n_nodes, const int32_t *parent, const int32_t *leaf_row, PairProgram &out);

// One step of the QUAD walk (models with K <= 4 states).  The tree is cut into "caps": connected pieces hanging below a
// node r whose state is known (the root, or a frontier node of an earlier cap), with at most 4 FRONTIER nodes (tree
// leaves, or internal nodes whose state is kept in a slot and that root further caps); every other node of the cap is
// MARGINALISED -- its state is summed out of the joint distribution of the frontier given r and never drawn.  One
// record = one cap = ONE draw from a 256-outcome alias row (o = sum_k state(out[k]) << 2k).  Only the leaves'
// joint distribution matters to simulateAndFlatten (MarkovProcessAlongTree.hs:71-98), and it is unchanged.
constexpr int QUAD_MAX_CAP = 12;  // nodes of a cap without its root (binary trees need at most 7)
struct QuadCapNode {
  int32_t node;   // preorder index (selects the P(t) table of the branch above it)
  int8_t par;     // index of its parent in cap[], -1: the cap's root r
  int8_t fr;      // frontier position 0..3, -1: marginalised
};
struct QuadRec {
  int32_t root;       // r
  int32_t out[4];     // frontier nodes, -1: absent; out[0] keys the record's Philox counters
  int16_t src;        // slot holding r's states
  int8_t dst[4];      // slot receiving out[k]'s states (internal frontier nodes), -1: none
  int32_t leaf[4];    // output row of out[k], -1: internal / absent
  int32_t n_cap;
  QuadCapNode cap[QUAD_MAX_CAP];  // in preorder of the cap (parents first)
};

struct QuadProgram {
  std::vector<QuadRec> recs;
  int n_slots = 0;
};

// Cuts the tree into the minimum number of caps (tree DP over the open cap's frontier size), orders the records depth
// first with the heaviest frontier node last and assigns slots (slot 0 = the root states; a record reads src before it
// writes dst, so a slot freed by a node's last cap is reused at once).  Expects arrays compile_walk accepted.
// max_frontier = 2 gives the caps of duo mode (protein: 20 x 20 outcomes per draw, elx_duo.cu).
void compile_quad_walk(int n_nodes, const int32_t *parent, const int32_t *leaf_row, QuadProgram &out, int max_frontier = 4);

// Symmetric eigen-decomposition (cyclic Jacobi) of S = D^{1/2} Q D^{-1/2}, Q = setDiagonal(E diag(pi)).
// Produces lambda[K], A = D^{-1/2} V and B = V^T D^{1/2} so that exp(Qt) = I + A diag(expm1(lambda t)) B.
bool eigen_reversible(int K, const double *pi, const double *exch, double *lambda, double *A, double *B,
                      std::string &msg);

}  // namespace elx
