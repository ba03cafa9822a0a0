// akugpu_phone_probs -- the reference tool aku/phone_probs.cc re-hosted on the GPU library.
// Same flags (aku/phone_probs.cc:60-81) and the same per-utterance LNA files; utterances of a recipe
// are batched into GPU calls (-C clusters / --eval-minc / --eval-ming included: the Gaussian-clustering
// approximation; -S speakers: per-speaker / per-utterance parameters of feature modules, e.g. a CMLLR lin_transform,
// and a speaker's global model-level `model cmllr` transform).
#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <fstream>
#include <map>
#include <sstream>
#include "akugpu.hh"

struct Utt { std::string audio, lna, speaker, utterance; double start_time, end_time; };

// stat() == 0 is the reference's test for "exists" (aku/phone_probs.cc:180-190): an empty file is skipped too
static bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }
// a file name inside single quotes of a /bin/sh command line: ' -> '\''
static std::string sh_quote(const std::string &s)
{
  std::string q = "'";
  for (size_t i = 0; i < s.size(); i++) { if (s[i] == '\'') q += "'\\''"; else q += s[i]; }
  return q + "'";
}

// Output names as io::Stream understands them (aku/io.cc:35-130): "-" = standard output, a leading '|' = pipe to a
// command, a trailing ".gz" = through gzip; anything else a plain file.
struct OutStream {
  FILE *fp; bool is_pipe; bool is_stdout;
  explicit OutStream(const std::string &name) : fp(NULL), is_pipe(false), is_stdout(false) {
    if (name == "-") { fp = stdout; is_stdout = true; }
    else if (!name.empty() && name[0] == '|') { fp = popen(name.c_str() + 1, "w"); is_pipe = true; }
    else if (name.size() >= 3 && name.compare(name.size() - 3, 3, ".gz") == 0) { fp = popen(("gzip > " + sh_quote(name)).c_str(), "w"); is_pipe = true; }
    else fp = fopen(name.c_str(), "wb");
    if (!fp) throw std::string("could not open ") + name + " for writing";
  }
  ~OutStream() { if (fp && is_pipe) pclose(fp); else if (fp && !is_stdout) fclose(fp); else if (fp) fflush(fp); }
};

int main(int argc, char **argv)
{
  std::string base, gk, mc, ph, cfg, recipe, outdir, clusters, speakers;
  double eval_minc = 0, eval_ming = 0.1;       // defaults of aku/phone_probs.cc:74-75
  int lnabytes = 2, batch = 0, bindex = 0, info = 0, device = 0, precision = AKUGPU_F32;
  bool raw_input = false, lna_by_audio = false, no_overwrite = false, no_norm = false, sort_recipe = false;
  long max_batch_samples = 64L << 20;
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
        printf("usage: akugpu_phone_probs [OPTION...]\n"
               "  -b, --base=BASENAME    base filename for model files\n  -g, --gk=FILE  -m, --mc=FILE  -p, --ph=FILE\n"
               "  -c, --config=FILE      feature configuration\n  -r, --recipe=FILE      recipe file\n"
               "  -a, --afname           name LNA files by the audio file\n      --sort-recipe      sort recipe lines by speaker, useful with adaptation\n  -o, --output-dir=DIR   base path for LNAs\n"
               "  -R, --raw-input        raw audio input\n      --lnabytes=INT     2 (default) or 4\n"
               "  -n, --no-overwrite     prevent overwriting existing files\n  -N, --no-normalization\n"
               "  -S, --speakers=FILE    speaker configuration file (feature-module parameters per speaker / utterance,\n"
               "                         global `model cmllr` transforms per speaker)\n"
               "  -C, --clusters=FILE    Gaussian clustering (.gcl)\n      --eval-minc=FLOAT  minimum ratio of top clusters to evaluate (0)\n"
               "      --eval-ming=FLOAT  minimum ratio of Gaussians to evaluate (0.1)\n"
               "  -B, --batch=INT  -I, --bindex=INT   recipe batching\n  -i, --info=INT\n"
               "      --precision=f32|f64   throughput (default) or parity arithmetic\n      --device=INT\n");
        return 0;
      } else if (a == "-b" || a == "--base") base = val();
      else if (a == "-g" || a == "--gk") gk = val();
      else if (a == "-m" || a == "--mc") mc = val();
      else if (a == "-p" || a == "--ph") ph = val();
      else if (a == "-c" || a == "--config") cfg = val();
      else if (a == "-r" || a == "--recipe") recipe = val();
      else if (a == "-o" || a == "--output-dir") outdir = val();
      else if (a == "-a" || a == "--afname" || a == "--lnabyaudio") lna_by_audio = true;
      else if (a == "--sort-recipe") sort_recipe = true;
      else if (a == "-R" || a == "--raw-input") raw_input = true;
      else if (a == "--lnabytes") lnabytes = atoi(val().c_str());
      else if (a == "-n" || a == "--no-overwrite") no_overwrite = true;
      else if (a == "-N" || a == "--no-normalization") no_norm = true;
      else if (a == "-B" || a == "--batch") batch = atoi(val().c_str());
      else if (a == "-I" || a == "--bindex") bindex = atoi(val().c_str());
      else if (a == "-i" || a == "--info") info = atoi(val().c_str());
      else if (a == "--precision") precision = (val() == "f64") ? AKUGPU_F64 : AKUGPU_F32;
      else if (a == "--device") device = atoi(val().c_str());
      else if (a == "-C" || a == "--clusters") clusters = val();
      else if (a == "--eval-minc") eval_minc = atof(val().c_str());
      else if (a == "--eval-ming") eval_ming = atof(val().c_str());
      else if (a == "-S" || a == "--speakers") speakers = val();
      else throw std::string("unknown option ") + a;
    }
    if (cfg.empty()) throw std::string("Must give --config");
    if (recipe.empty()) throw std::string("Must give --recipe");
    if (lnabytes != 2 && lnabytes != 4) throw std::string("Invalid number of LNA bytes");   // aku/phone_probs.cc:88-90
    if (!outdir.empty() && outdir[outdir.size() - 1] != '/') outdir += "/";

    akugpu::Engine eng(device);
    akugpu::FeatureGenerator gen(eng);
    gen.load_configuration(cfg);
    akugpu::HmmSet model(eng);
    if (!base.empty()) model.read_all(base);
    else if (!gk.empty() && !mc.empty() && !ph.empty()) model.read_files(gk, mc, ph);
    else throw std::string("Must give either --base or all --gk, --mc and --ph");
    akugpu::SpeakerConfig speaker_conf(eng);
    if (!speakers.empty()) speaker_conf.read_speaker_file(speakers);     // aku/phone_probs.cc:93-94,133-134
    if (!clusters.empty()) {          // aku/phone_probs.cc:112-117
      model.read_clustering(clusters);
      model.set_clustering_min_evals(eval_minc, eval_ming);
    }
    if (model.dim() != gen.dim()) {   // aku/phone_probs.cc:119-124
      char msg[256];
      snprintf(msg, sizeof msg, "Gaussian dimension is %d but feature dimension is %d.", model.dim(), gen.dim());
      throw std::string(msg);
    }
    const int S = model.num_states();
    if ((batch != 0) != (bindex != 0)) throw std::string("Must give both --batch and --bindex");   // aku/phone_probs.cc:135-136
    akugpu::Recipe rcp;
    rcp.read(recipe, batch, bindex);
    if (sort_recipe) rcp.sort_infos();                                                                // aku/phone_probs.cc:141-142
    std::vector<Utt> utts;
    for (size_t k = 0; k < rcp.infos.size(); k++) {
      const akugpu::Recipe::Info &r = rcp.infos[k];
      Utt u;
      u.audio = r.audio_path; u.lna = r.lna_path; u.speaker = r.speaker_id; u.utterance = r.utterance_id;
      u.start_time = r.start_time; u.end_time = r.end_time;
      utts.push_back(u);
    }
    // output names (aku/phone_probs.cc:155-176) and --no-overwrite (:180-190)
    std::vector<Utt> todo;
    for (size_t k = 0; k < utts.size(); k++) {
      Utt u = utts[k];
      if (lna_by_audio) {               // aku/phone_probs.cc:158-176
        std::string f = u.audio;
        size_t sl = f.rfind('/'); if (sl != std::string::npos && sl + 1 < f.size()) f = f.substr(sl + 1);   // a trailing '/' keeps the path
        size_t dot = f.rfind('.'); if (dot != std::string::npos && dot > 0) f.erase(dot);                    // ".hidden" keeps its name
        u.lna = f + ".lna";
      }
      u.lna = outdir + u.lna;            // an empty lna= is an open error later, as in the reference (no silent fallback)
      if (no_overwrite && file_exists(u.lna)) {
        fprintf(stderr, "WARNING: skipping existing lna file %s\n", u.lna.c_str());
        continue;
      }
      todo.push_back(u);
    }
    const float fr = gen.frame_rate();
    size_t i = 0;
    std::vector<uint8_t> rec;
    while (i < todo.size()) {
      std::vector<int16_t> pcm;
      std::vector<int64_t> uo(1, 0);
      size_t j = i;
      if (!speakers.empty()) {        // aku/phone_probs.cc:192-197; a GPU call covers utterances that share the parameters
        speaker_conf.set_speaker(todo[i].speaker);
        if (!todo[i].utterance.empty()) speaker_conf.set_utterance(todo[i].utterance);
      }
      while (j < todo.size() && (j == i || ((long)pcm.size() < max_batch_samples &&
                                            (speakers.empty() || (todo[j].speaker == todo[i].speaker && todo[j].utterance == todo[i].utterance))))) {
        std::vector<int16_t> one;
        int rate = 0;
        akugpu::read_audio(todo[j].audio, gen.sample_rate(), raw_input || gen.config_raw(), one, rate, gen.config_big_endian());
        if (rate != gen.sample_rate()) {
          char msg[256];
          snprintf(msg, sizeof msg, "Audio file sample rate (%d Hz) and model configuration (%d Hz) don't agree.", rate, gen.sample_rate());
          throw std::string(msg);
        }
        pcm.insert(pcm.end(), one.begin(), one.end());
        uo.push_back((int64_t)pcm.size());
        j++;
      }
      const int n = (int)(j - i);
      std::vector<int64_t> fo(n + 1, 0);
      akugpu::check(eng.ctx(), akugpu_features(eng.ctx(), NULL, uo.data(), n, NULL, 0, fo.data()));
      rec.resize((size_t)fo[n] * S * lnabytes);
      akugpu::check(eng.ctx(), akugpu_phone_probs(eng.ctx(), pcm.data(), uo.data(), n, precision, lnabytes, no_norm ? 0 : 1,
                                                  rec.data(), fo.data(), NULL));
      for (int k = 0; k < n; k++) {
        const Utt &u = todo[i + k];
        if (info > 0) printf("Processing file: %s\n", u.audio.c_str());
        const int64_t nf = fo[k + 1] - fo[k];
        int64_t s = (int64_t)(int)(u.start_time * fr), e = (int64_t)(int)(u.end_time * fr);   // aku/phone_probs.cc:199-206
        if (e == 0 || e > nf) e = nf;          // the reference's loop breaks at eof
        if (s > e) s = e;
        OutStream os(u.lna);
        FILE *fp = os.fp;
        uint8_t hdr[5];
        akugpu_lna_header(S, lnabytes, hdr);
        if (fwrite(hdr, 1, 5, fp) != 5) throw std::string("Write error");
        if (s < 0) {
          // frames before the file: the reference's loop starts at a negative frame and every module sees the
          // border-replicated first window there (aku/FeatureModules.cc:381-422); generated on their own
          const int64_t nneg = std::min<int64_t>(e, 0) - s;
          const int dim = gen.dim();
          std::vector<double> fneg((size_t)nneg * dim);
          int d_out = 0;
          akugpu::check(eng.ctx(), akugpu_features_range(eng.ctx(), pcm.data() + uo[k], uo[k + 1] - uo[k], (int)s, (int)(s + nneg), NULL,
                                                         fneg.data(), 1, &d_out));
          std::vector<uint8_t> rneg((size_t)nneg * S * lnabytes);
          akugpu::check(eng.ctx(), akugpu_gmm_lna(eng.ctx(), fneg.data(), 1, nneg, precision, lnabytes, no_norm ? 0 : 1, rneg.data()));
          if (fwrite(rneg.data(), 1, rneg.size(), fp) != rneg.size()) throw std::string("Write error");
          s = std::min<int64_t>(e, 0);
        }
        const size_t nbytes = (size_t)(e - s) * S * lnabytes;
        if (nbytes && fwrite(rec.data() + (size_t)(fo[k] + s) * S * lnabytes, 1, nbytes, fp) != nbytes)
          throw std::string("Write error");
      }
      i = j;
    }
  } catch (std::string &str) {
    fprintf(stderr, "exception: %s\n", str.c_str());
    return 1;
  } catch (std::exception &e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
