// route_graph_check.cu -- TEST HARNESS ONLY (host code, no device needed): the memoised catchment search of
// vicres_route.cu (catchment_maps) against the path-by-path search it replaces (list_cells / reaches, the literal form of
// SEARCH_WHOLECATCHMENT / SEARCH_CATCHMENTRS) on random grids with cycles, dead ends and paths that leave the grid.
#include "../../vicres_b200/csrc/vicres_route.cu"
#include <random>
int main()
{
  std::mt19937 rng(5);
  int bad = 0, total = 0;
  for (int trial = 0; trial < 60; trial++) {
    const int nrow = 5 + rng() % 20, ncol = 5 + rng() % 20, n = nrow * ncol;
    std::vector<int32_t> dir(n), reser(n, 0);
    std::vector<float> f(n, 1.f);
    for (int k = 0; k < n; k++) dir[k] = rng() % 9;                     // random directions incl. 0 (no flow), cycles happen
    for (int k = 0; k < n; k++) { int r = rng() % 12; reser[k] = r == 0 ? 1 + (int)(rng() % 50) : (r == 1 ? 9999 : 0); }
    vr_route_grid g;
    memset(&g, 0, sizeof g);
    g.ncol = ncol; g.nrow = nrow; g.dir = dir.data(); g.reser = reser.data(); g.velo = f.data(); g.diff = f.data(); g.xmask = f.data();
    Graph G;
    build_graph(&g, G);
    const int station = rng() % n;
    CatchmentMaps cm;
    catchment_maps(G, station, cm);
    std::vector<int> ref, mine;
    list_cells(G, station, 0, ref);
    for (int I = 1; I <= ncol; I++) for (int J = 1; J <= nrow; J++) { int k = gidx(G, I, J); if (cm.whole[k]) mine.push_back(k); }
    total++; if (ref != mine) { bad++; printf("whole differs trial %d\n", trial); }
    list_cells(G, station, 2, ref); mine.clear();
    for (int I = 1; I <= ncol; I++) for (int J = 1; J <= nrow; J++) { int k = gidx(G, I, J); if (cm.direct[k]) mine.push_back(k); }
    total++; if (ref != mine) { bad++; printf("direct differs trial %d\n", trial); }
    for (int T = 0; T < n; T++) {
      if (!(G.reser[T] > 0 && G.reser[T] != 9999)) continue;
      list_cells(G, T, 1, ref); mine.clear();
      for (int I = 1; I <= ncol; I++) for (int J = 1; J <= nrow; J++) { int k = gidx(G, I, J); if (cm.first_res[k] == T) mine.push_back(k); }
      total++; if (ref != mine) { bad++; if (bad < 10) printf("reservoir list differs trial %d T %d (%zu vs %zu)\n", trial, T, ref.size(), mine.size()); }
    }
  }
  printf("%d of %d lists differ\n", bad, total);
  return bad != 0;
}
