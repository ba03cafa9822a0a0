// akugpu_feacat -- the reference tool aku/feacat.cc re-hosted on the GPU library: prints the feature
// vectors of one audio (or `pre` feature) file.  Same flags (aku/feacat.cc:50-62) and output formats
// (ASCII "%8.4f " rows, or raw float32 with an optional int32 dimension header, aku/feacat.cc:15-33,88-95);
// the requested frame range is computed in one GPU call instead of frame by frame.
// Not carried over: -G (Gaussian noise from the reference's ziggurat generator; a test aid, refused here).
#include <limits.h>
#include <stdlib.h>
#include <fstream>
#include <sstream>
#include "akugpu.hh"

static void print_feature(const double *fea, int dim, bool raw_output)
{
  if (raw_output) {
    for (int i = 0; i < dim; i++) {
      float tmp = (float)fea[i];
      if (fwrite(&tmp, sizeof(float), 1, stdout) != 1) throw std::string("write failed");
    }
  } else {
    for (int i = 0; i < dim; i++) printf("%8.4f ", fea[i]);
    printf("\n");
  }
}

int main(int argc, char **argv)
{
  std::string cfg, write_cfg, speakers, speaker_id, utterance_id;
  std::vector<std::string> args;
  bool raw_output = false, header = false, have_start = false, have_end = false, have_utt = false;
  int start_frame = 0, end_frame = INT_MAX, device = 0;
  try {
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i], v;
      size_t eq = a.find('=');
      if (a.compare(0, 2, "--") == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); }
      auto val = [&]() -> std::string {
        if (!v.empty()) return v;
        if (i + 1 >= argc) throw std::string("missing value for ") + a;
        return argv[++i];
      };
      if (a == "-h" || a == "--help") {
        printf("usage: akugpu_feacat [OPTION...] FILE\n"
               "  -c, --config=FILE        read feature configuration\n  -w, --write-config=FILE  write feature configuration\n"
               "      --raw-output         raw float output\n  -H, --header             write a header (feature dim, 32 bits) in raw output\n"
               "  -s, --start-frame=INT    audio start frame\n  -e, --end-frame=INT      audio end frame (inclusive)\n"
               "  -S, --speakers=FILE      speaker configuration file\n  -d, --speaker-id=NAME    speaker ID\n"
               "  -u, --utterance-id=NAME  utterance ID\n      --device=INT\n");
        return 0;
      } else if (a == "-c" || a == "--config") cfg = val();
      else if (a == "-w" || a == "--write-config") write_cfg = val();
      else if (a == "--raw-output") raw_output = true;
      else if (a == "-H" || a == "--header") header = true;
      else if (a == "-s" || a == "--start-frame") { start_frame = atoi(val().c_str()); have_start = true; }
      else if (a == "-e" || a == "--end-frame") { end_frame = atoi(val().c_str()); have_end = true; }
      else if (a == "-S" || a == "--speakers") speakers = val();
      else if (a == "-d" || a == "--speaker-id") speaker_id = val();
      else if (a == "-u" || a == "--utterance-id") { utterance_id = val(); have_utt = true; }
      else if (a == "-G" || a == "--gaussian-std") throw std::string("-G (feature noise) is not supported by akugpu_feacat");
      else if (a == "--device") device = atoi(val().c_str());
      else if (a == "-" || a[0] != '-') args.push_back(a);
      else throw std::string("unknown option ") + a;
    }
    (void)have_start; (void)have_end;
    if (args.size() != 1) { fprintf(stderr, "usage: akugpu_feacat [OPTION...] FILE\n"); return 1; }
    if (cfg.empty()) throw std::string("Must give --config");
    if (header && !raw_output) fprintf(stderr, "Warning: header is only written in raw output mode\n");

    akugpu::Engine eng(device);
    akugpu::FeatureGenerator gen(eng);
    gen.load_configuration(cfg);
    // The reference sets the speaker after opening the file (aku/feacat.cc:74-84); its modules are lazy, so the order
    // does not matter there.  Here open() already computes the file's frames, hence parameters first.
    akugpu::SpeakerConfig speaker_conf(eng);
    if (!speakers.empty()) {
      speaker_conf.read_speaker_file(speakers);
      speaker_conf.set_speaker(speaker_id);
      if (have_utt) speaker_conf.set_utterance(utterance_id);
    }
    gen.open(args[0]);

    if (!write_cfg.empty()) {
      // FeatureGenerator::write_configuration (aku/FeatureGenerator.cc:222-246) re-serialises the module list; the
      // configuration text read is an equivalent serialisation of the same chain, so it is written back as it is.
      std::ifstream in(cfg.c_str());
      std::ofstream out(write_cfg.c_str());
      if (!in || !out) throw std::string("could not write configuration ") + write_cfg;
      out << in.rdbuf();
    }

    const int dim = gen.dim();
    if (raw_output && header && fwrite(&dim, sizeof(int), 1, stdout) != 1) throw std::string("write failed");

    std::vector<double> rows;
    if (start_frame < end_frame) {                    // ascending, end inclusive; open end = to the end of the file
      const int last = end_frame == INT_MAX ? gen.num_frames() - 1 : end_frame;
      if (start_frame >= 0 && last < gen.num_frames()) {
        for (int f = start_frame; f <= last; f++) print_feature(&gen.features()[(size_t)f * dim], dim, raw_output);
      } else if (last >= start_frame) {
        gen.generate_range(start_frame, last + 1, rows);
        for (int f = start_frame; f <= last; f++) print_feature(&rows[(size_t)(f - start_frame) * dim], dim, raw_output);
      }
    } else {                                          // descending (aku/feacat.cc:112-120)
      gen.generate_range(end_frame, start_frame + 1, rows);
      for (int f = start_frame; f >= end_frame; f--) print_feature(&rows[(size_t)(f - end_frame) * dim], dim, raw_output);
    }
    fflush(stdout);
    gen.close();
  } catch (std::string &str) {
    fprintf(stderr, "exception: %s\n", str.c_str());
    return 1;
  }
  return 0;
}
