/*
 * elynx_b200 -- host-side helpers above the simulator ABI (include/elynx_b200.h).
 *
 * In a Haskell drop-in these stay Haskell (the model algebra and parsers of elynx-markov / elynx-tree /
 * slynx are untouched by the GPU path).  They exist here, in C++ behind a C ABI, because the image has no
 * GHC: the thin CLI (`python -m elynx_b200.cli simulate`), bench.py and the tests need to build the same
 * (weights, pi, exch) triples and trees the reference's caller would hand to the simulator.
 * Nothing in this header touches the GPU.
 *
 * Reference functions restated (paths relative to /root/reference):
 *   model string grammar   slynx/src/SLynx/Simulate/PhyloModel.hs:73-231 (getPhyloModel)
 *   substitution models    elynx-markov/src/ELynx/MarkovProcess/{Nucleotide.hs:59-104,AminoAcid.hs:46-646}
 *   normalisation          elynx-markov/src/ELynx/MarkovProcess/{RateMatrix.hs:57-103,SubstitutionModel.hs:101-130,MixtureModel.hs:67-115}
 *   gamma expansion        elynx-markov/src/ELynx/MarkovProcess/GammaRateHeterogeneity.hs:47-123
 *   CXX profile mixtures   elynx-markov/src/ELynx/MarkovProcess/CXXModels.hs:31-148
 *   EDM / site profiles    elynx-markov/src/ELynx/Import/MarkovProcess/{EDMModelPhylobayes.hs:33-64,SiteprofilesPhylobayes.hs:33-62}
 *   Newick (Standard)      elynx-tree/src/ELynx/Tree/Import/Newick.hs:90-201 + Phylogeny.hs:374-381 (toLengthTree)
 *   birth-death trees      elynx-tree/src/ELynx/Tree/Simulate/PointProcess.hs:97-317 (tlynx simulate)
 *
 * All functions return 0 or a negative elx_status; elx_last_error(NULL) holds the message.
 */
#ifndef ELYNX_B200_MODEL_H
#define ELYNX_B200_MODEL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct elx_phylo_model elx_phylo_model; /* PhyloModel = SubstitutionModel | MixtureModel (PhyloModel.hs:30-31) */
typedef struct elx_flat_tree elx_flat_tree;     /* preorder flattening of `Tree Length Name` */

#define ELX_ALPHABET_DNA 0
#define ELX_ALPHABET_PROTEIN 1

/* getPhyloModel (PhyloModel.hs:211-231).  subst: `-s` string or NULL; mix: `-m` string or NULL;
 * global_norm: `-n`; weights/n_weights: `-w` or NULL/0; EDM components (`-e` / `-p` files, already parsed):
 * n_edm components of edm_k frequencies each, or NULL/0/0. */
int elx_phylo_model_parse(const char *subst, const char *mix, int global_norm, const double *weights, int n_weights,
                          const double *edm_weights, const double *edm_pis, int n_edm, int edm_k, elx_phylo_model **out);
/* expand n alpha (GammaRateHeterogeneity.hs:47-51): turns the model into a mixture in place. */
int elx_phylo_model_expand_gamma(elx_phylo_model *m, int n_categories, double alpha);
void elx_phylo_model_free(elx_phylo_model *m);
int elx_phylo_model_info(const elx_phylo_model *m, int *n_states, int *n_components, int *is_mixture, int *alphabet);
/* weights [C], pi [C][K], exch [C][K][K] -- exactly what simulateAlignment extracts (slynx Simulate.hs:104-116) */
int elx_phylo_model_arrays(const elx_phylo_model *m, double *weights, double *pi, double *exch);
/* name of component c (or of the model itself for c < 0); valid until the model is freed */
const char *elx_phylo_model_name(const elx_phylo_model *m, int c);

/* getMeans n alpha (GammaRateHeterogeneity.hs:88-107), closed form */
int elx_gamma_means(int n, double alpha, double *out);

/* Phylobayes EDM file (EDMModelPhylobayes.hs:33-64) and site-profile file (SiteprofilesPhylobayes.hs:33-62).
 * On success *weights ([n]) and *pis ([n][k]) are malloc'ed; release with elx_free. */
int elx_parse_edm_phylobayes(const char *text, int64_t len, int *n_components, int *k, double **weights, double **pis);
int elx_parse_siteprofiles_phylobayes(const char *text, int64_t len, int *n_components, int *k, double **weights, double **pis);
void elx_free(void *p);

/* newick Standard (Newick.hs:90-96) followed by toLengthTree (Phylogeny.hs:374-381) */
int elx_newick_parse(const char *text, int64_t len, elx_flat_tree **out);
/* simulateReconstructedTree n Random lambda mu (PointProcess.hs:119-148,151-177,245-317), sub-critical or critical
 * process, leaf names "0".."n-1"; deterministic in `seed` (own SplitMix64 stream, not the reference's) */
int elx_birth_death_tree(int n_leaves, double lambda, double mu, uint64_t seed, elx_flat_tree **out);
void elx_flat_tree_free(elx_flat_tree *t);
int elx_flat_tree_size(const elx_flat_tree *t, int *n_nodes, int *n_leaves);
int elx_flat_tree_arrays(const elx_flat_tree *t, int32_t *parent, double *brlen, int32_t *leaf_row);
const char *elx_flat_tree_name(const elx_flat_tree *t, int node);
/* toNewick (elynx-tree Export/Newick.hs:49-65): returns a malloc'ed NUL-terminated string (elx_free) */
int elx_flat_tree_to_newick(const elx_flat_tree *t, char **out);

/* writeGZFile (elynx-tools/src/ELynx/Tools/InputOutput.hs:80-83) for big outputs: gzip `data` into `path` with n_threads
 * worker threads (<= 0: all host cores).  The file is a sequence of gzip members, one per block_bytes of input (<= 0: 4 MiB),
 * which gunzip, Python's gzip module and the reference's `decompress` (InputOutput.hs:73-76) read back as one stream.
 * level: zlib 0..9 (anything else: 6).  Host-only. */
int elx_gzip_write(const char *path, const uint8_t *data, int64_t size, int32_t n_threads, int32_t level, int64_t block_bytes);

#ifdef __cplusplus
}
#endif
#endif
