// host_caller.cpp -- stand-in for the Fortran caller: holds every field exactly as CICE would
// (a(nx_block,ny_block,max_blocks), column major; Fortran logicals as 4-byte ints with -1 for .true.
// to prove the mask conversion does not depend on the compiler's representation), calls the C++ host
// mirror of the reference seam, and writes the inout arrays back.
//   usage: host_caller case.bin out.bin
// case.bin: int32 header {nx_block, ny_block, nblocks, max_blocks, nx_global, ny_global, ew, ns, ndte, mode, kernel},
//           12 float64 scalars (arlx1i denom1 revp brlx e_factor epp2i capping Ktens u0 cosw sinw rhow),
//           int32 ilo ihi jlo jhi [nblocks each], i_glob [nx_block*nblocks], j_glob [ny_block*nblocks],
//           10 geometry arrays, 30 field arrays (struct order), 2 int32 masks.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dyn_evp_b200.hpp"

template <class T>
static std::vector<T> rd(FILE *f, size_t n) {
  std::vector<T> v(n);
  if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  auto h = rd<int32_t>(f, 11);
  auto sc = rd<double>(f, 12);
  const int nxb = h[0], nyb = h[1], nb = h[2], mb = h[3];
  const size_t n = (size_t)nxb * nyb * mb;
  auto ilo = rd<int32_t>(f, nb), ihi = rd<int32_t>(f, nb), jlo = rd<int32_t>(f, nb), jhi = rd<int32_t>(f, nb);
  auto ig = rd<int32_t>(f, (size_t)nxb * nb), jg = rd<int32_t>(f, (size_t)nyb * nb);
  std::vector<std::vector<double>> geo, fld;
  for (int q = 0; q < 10; ++q) geo.push_back(rd<double>(f, n));
  for (int q = 0; q < 30; ++q) fld.push_back(rd<double>(f, n));
  auto mT = rd<int32_t>(f, n), mU = rd<int32_t>(f, n);
  fclose(f);
  for (auto &m : mT) m = m ? -1 : 0;  // an ifort-style .true.
  for (auto &m : mU) m = m ? -1 : 0;

  static const char *bn[4] = {"open", "closed", "cyclic", "tripole"};
  cice_b200::BlockTable bt;
  bt.nx_block = nxb; bt.ny_block = nyb; bt.nblocks = nb; bt.max_blocks = mb;
  bt.nx_global = h[4]; bt.ny_global = h[5]; bt.ew_boundary_type = bn[h[6]]; bt.ns_boundary_type = bn[h[7]];
  bt.ilo = ilo.data(); bt.ihi = ihi.data(); bt.jlo = jlo.data(); bt.jhi = jhi.data(); bt.i_glob = ig.data(); bt.j_glob = jg.data();
  cice_b200::Geometry g{geo[0].data(), geo[1].data(), geo[2].data(), geo[3].data(), geo[4].data(),
                        geo[5].data(), geo[6].data(), geo[7].data(), geo[8].data(), geo[9].data()};
  cice_b200::EvpScalars s;
  s.ndte = h[8]; s.mode = h[9]; s.kernel = h[10];
  s.arlx1i = sc[0]; s.denom1 = sc[1]; s.revp = sc[2]; s.brlx = sc[3]; s.e_factor = sc[4]; s.epp2i = sc[5];
  s.capping = sc[6]; s.Ktens = sc[7]; s.u0 = sc[8]; s.cosw = sc[9]; s.sinw = sc[10]; s.rhow = sc[11];
  try {
    // calling run before init must abort the way the reference would
    bool aborted = false;
    try {
      cice_b200::dyn_evp_b200_run(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s);
    } catch (const cice_b200::AbortIce &) { aborted = true; }
    if (!aborted) { fprintf(stderr, "run before init did not abort\n"); return 3; }

    cice_b200::dyn_evp_b200_init(bt, g);
    auto D = [&](int q) { return fld[q].data(); };
    // struct order: 12 stresses, strength, cdn, aiU, uocn, vocn, waterx, watery, forcex, forcey, umassdti, fmU,
    //               strintx, strinty, TbU, taubx, tauby, uvel, vvel
    cice_b200::dyn_evp_b200_run(D(0), D(1), D(2), D(3), D(4), D(5), D(6), D(7), D(8), D(9), D(10), D(11), D(12), D(13), D(14), D(15),
                                D(16), D(17), D(18), D(19), D(20), D(21), D(22), D(23), D(24), D(25), D(26), D(27), D(28), D(29),
                                mT.data(), mU.data(), s);
    cice_b200::dyn_evp_b200_finalize();
  } catch (const cice_b200::AbortIce &e) {
    fprintf(stderr, "abort_ice: %s\n", e.what());
    return 128;  // MPI_ABORT(comm,128) in the reference
  }
  FILE *o = fopen(argv[2], "wb");
  for (int q = 0; q < 30; ++q) fwrite(fld[q].data(), sizeof(double), n, o);
  fclose(o);
  printf("host_caller ok\n");
  return 0;
}
