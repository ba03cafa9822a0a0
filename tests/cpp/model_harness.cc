// CPU harness for the library's OWN model-file and clustering-file readers (aaltoasr_b200/csrc/model.cu:
// akugpu::model_read_files, akugpu::model_read_clustering -- pure host code, exported from libakugpu.so).  They need no
// device, so tests/test_abi.py runs them here against the arrays the files were written from, against the reference's
// HmmSet on the same files, and on malformed files (no crash, an error that names the problem).
//   model_harness read BASE OUT.bin            dumps S G D n_full, offsets, indices, weights, means, covs, full covs
//   model_harness gcl  BASE FILE.gcl OUT.bin   dumps n_clusters, gauss_cluster[G], sizes, centre means / covs
//   model_harness expand BASE OUT.bin          akugpu::tc_expanded_params (csrc/gmm_tc.cu): centre, theta, gconst, q per Gaussian
//   model_harness slots BASE GROUP OUT.bin     akugpu::tc_build_slots: slot -> state / first component / flags
#include <stdio.h>
#include <string>
#include "../../aaltoasr_b200/csrc/ctx.hpp"
#include "../../aaltoasr_b200/csrc/kernels.hpp"

namespace akugpu {      // host helpers of the tensor-core packers (declared in csrc/tc_common.cuh, which is device code)
double tc_expanded_params(const HostModel &hm, bool full, int L, std::vector<double> &cen, std::vector<double> &theta,
                          std::vector<double> &gconst, std::vector<double> *q_of_gauss);
void tc_build_slots(const HostModel &hm, int group, std::vector<int> &slot_state, std::vector<int> &slot_k0, std::vector<int> &slot_flags,
                    const std::vector<char> *skip_state);
}

using namespace akugpu;

template <class T>
static void put(FILE *fp, const std::vector<T> &v)
{
  long long n = (long long)v.size();
  fwrite(&n, sizeof n, 1, fp);
  if (n) fwrite(v.data(), sizeof(T), v.size(), fp);
}

int main(int argc, char **argv)
{
  if (argc < 4) return 2;
  const std::string mode = argv[1], base = argv[2];
  try {
    HostModel hm;
    model_read_files(base + ".gk", base + ".mc", base + ".ph", hm);
    FILE *fp = fopen(argv[argc - 1], "wb");
    if (!fp) return 3;
    if (mode == "read") {
      std::vector<int32_t> hdr = {hm.S, hm.G, hm.D, hm.n_full};
      put(fp, hdr); put(fp, hm.mix_off); put(fp, hm.mix_gauss); put(fp, hm.mix_w); put(fp, hm.mean); put(fp, hm.cov);
      put(fp, hm.full_index); put(fp, hm.full_cov);
    } else if (mode == "expand") {
      const bool full = hm.n_full > 0;
      const int L = full ? hm.D * (hm.D + 3) / 2 : 2 * hm.D;
      std::vector<double> cen, theta, gconst, q, qmax(1);
      qmax[0] = tc_expanded_params(hm, full, L, cen, theta, gconst, &q);
      put(fp, cen); put(fp, theta); put(fp, gconst); put(fp, q); put(fp, qmax);
    } else if (mode == "slots") {
      std::vector<int> st, k0, fl;
      std::vector<char> skip(hm.S, 0);
      for (int s = 0; s < hm.S; s += 5) skip[s] = 1;                   // every fifth state left to another kernel
      tc_build_slots(hm, atoi(argv[3]), st, k0, fl, nullptr);
      put(fp, st); put(fp, k0); put(fp, fl);
      tc_build_slots(hm, atoi(argv[3]), st, k0, fl, &skip);
      put(fp, st); put(fp, k0); put(fp, fl);
    } else {
      model_read_clustering(argv[3], hm);
      std::vector<int32_t> hdr = {hm.n_clusters}, sizes;
      for (size_t c = 0; c < hm.cluster_gauss.size(); c++) sizes.push_back((int32_t)hm.cluster_gauss[c].size());
      put(fp, hdr); put(fp, hm.gauss_cluster); put(fp, sizes); put(fp, hm.c_mean); put(fp, hm.c_cov);
    }
    fclose(fp);
  } catch (Error &e) {
    printf("error %d: %s\n", e.code, e.msg.c_str());
    return 1;
  }
  return 0;
}
