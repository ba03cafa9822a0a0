// Driver for akugpu::PPToolbox (aaltoasr_b200/csrc/host/akugpu.hh), the C++ mirror of aku::PPToolbox
// (aku/PhoneProbsToolbox.hh).  tests/test_gpu_host.py compiles it against libakugpu.so and compares its LNA files
// with the library's own output and the reference's; tests/test_abi.py checks that it builds and fails loudly
// without a CUDA device.
//   pptoolbox_main CFG MODEL_BASE IN.wav OUT.lna [f32|f64] [file|fd|rawfd] [CLUSTERS.gcl MINC MING]
#include "../../aaltoasr_b200/csrc/host/akugpu.hh"

int main(int argc, char **argv)
{
  if (argc < 5) { fprintf(stderr, "usage: pptoolbox_main CFG BASE IN OUT [f32|f64] [file|fd|rawfd] [GCL MINC MING]\n"); return 2; }
  const std::string prec = argc > 5 ? argv[5] : "f32", mode = argc > 6 ? argv[6] : "file";
  try {
    akugpu::PPToolbox pp(0, prec == "f64" ? AKUGPU_F64 : AKUGPU_F32);
    pp.read_configuration(argv[1]);
    pp.read_models(argv[2]);
    if (argc > 9) pp.set_clustering(argv[7], atof(argv[8]), atof(argv[9]));
    if (mode == "file") {
      pp.generate(argv[3], argv[4], false);
    } else {
      int in = open(argv[3], O_RDONLY), out = open(argv[4], O_WRONLY | O_CREAT | O_TRUNC, 0664);
      if (in < 0 || out < 0) throw std::string("could not open the files");
      pp.generate_to_fd(in, out, mode == "rawfd");
      close(in);
      close(out);
    }
  } catch (std::string &s) {
    fprintf(stderr, "exception: %s\n", s.c_str());
    return 1;
  }
  return 0;
}
