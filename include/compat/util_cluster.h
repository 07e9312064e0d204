// include/compat/util_cluster.h -- the reference's clustering entry point (src/util_cluster.h:75, used by
// src/BreakID.cc:1304-1326) on top of the B200 library.
//
// init_cluster() runs the whole agglomerative clustering of one point set on the GPU (bkid_op_cluster: distance cut,
// connected components, exact replay of the reference's merge order) and hands back the ROOT nodes only: the callers in
// the reference (add_cluster_id_for_enspan_vec, print_root_nodes) look at nothing else.  Root nodes keep the reference's
// node order -- unmerged leaves first, in point order, then merged roots in creation order -- so numbering clusters by
// walking `nodes` gives the reference's cluster ids.  The dendrogram below the roots, the neighbour lists and the
// distance matrix are not materialised (distance_matrix is NULL, node::neighbours is NULL).
//
// Only linkage type 1 (the one src/BreakID.cc:32 uses) is implemented; init_cluster throws std::invalid_argument for others.
#pragma once
#include <string>
#include <vector>

#define NOT_USED 0
#define LEAF_NODE 1
#define MERGER 2

struct coordinate { double x, y; };
struct point { coordinate pos; std::string label; int cluster_id = -1; };
struct neighbour { int target; double distance; neighbour *prev, *next; };
struct node {
  int type = NOT_USED, is_root = 0, height = 0;
  coordinate centroid{0, 0};
  std::string label;
  std::vector<int> merged;            // not filled
  int num_points = 0;
  std::vector<int> points;            // indices into the caller's point vector
  neighbour *neighbours = nullptr;    // not filled
};
struct cluster_struct {
  unsigned long num_points = 0;
  int num_root_clusters = 0;
  int num_nodes = 0;
  std::vector<node> nodes;            // root nodes only, reference node order
  double **distance_matrix = nullptr; // not materialised
};

void init_cluster(cluster_struct &main_cluster, long distance_threshold, std::vector<point> &points, int linkage_type);
int print_root_nodes(cluster_struct &main_cluster);     // number of root nodes (src/util_cluster.cc:398-417)
double euclidean_distance(coordinate &a, coordinate &b);
