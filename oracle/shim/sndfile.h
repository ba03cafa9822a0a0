// Stand-in for libsndfile's <sndfile.h> (not installed in this image).
// TEST INFRASTRUCTURE ONLY: lets the reference's aku/AudioReader.cc compile
// unmodified for the oracle binaries under oracle/_ref/.  Supports what that
// file calls (aku/AudioReader.cc:92-137,164,197,255): RIFF/WAVE PCM16 and
// headerless RAW PCM16 (either endianness), read fully into memory.
#ifndef ORACLE_SHIM_SNDFILE_H
#define ORACLE_SHIM_SNDFILE_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <vector>

typedef int64_t sf_count_t;
struct SF_INFO { sf_count_t frames; int samplerate, channels, format, sections, seekable; };
enum { SFM_READ = 0x10 };
enum {
  SF_FORMAT_WAV = 0x010000, SF_FORMAT_RAW = 0x040000, SF_FORMAT_PCM_16 = 0x0002,
  SF_FORMAT_SUBMASK = 0x0000FFFF, SF_FORMAT_TYPEMASK = 0x0FFF0000,
  SF_ENDIAN_FILE = 0x00000000, SF_ENDIAN_LITTLE = 0x10000000, SF_ENDIAN_BIG = 0x20000000
};
struct SNDFILE_tag { std::vector<short> pcm; sf_count_t pos; int fd; int close_fd; };
typedef struct SNDFILE_tag SNDFILE;

static inline SNDFILE *shim_sf_from_bytes(std::vector<unsigned char> &b, SF_INFO *info) {
  SNDFILE *s = new SNDFILE_tag; s->pos = 0; s->fd = -1; s->close_fd = 0;
  bool raw = (info->format & SF_FORMAT_TYPEMASK) == SF_FORMAT_RAW;
  size_t data_off = 0, data_len = b.size();
  bool big = false;
  if (raw) {
    big = (info->format & SF_ENDIAN_BIG) != 0;
  } else {
    if (b.size() < 12 || memcmp(&b[0], "RIFF", 4) || memcmp(&b[8], "WAVE", 4)) { delete s; return NULL; }
    size_t p = 12; int fmt_tag = -1, bits = 0, ch = 0, sr = 0; bool have_data = false;
    while (p + 8 <= b.size()) {
      uint32_t len = b[p + 4] | (b[p + 5] << 8) | (b[p + 6] << 16) | ((uint32_t)b[p + 7] << 24);
      if (!memcmp(&b[p], "fmt ", 4) && p + 8 + 16 <= b.size()) {
        fmt_tag = b[p + 8] | (b[p + 9] << 8); ch = b[p + 10] | (b[p + 11] << 8);
        sr = b[p + 12] | (b[p + 13] << 8) | (b[p + 14] << 16) | ((uint32_t)b[p + 15] << 24);
        bits = b[p + 22] | (b[p + 23] << 8);
      } else if (!memcmp(&b[p], "data", 4)) {
        data_off = p + 8; data_len = len; if (data_off + data_len > b.size()) data_len = b.size() - data_off;
        have_data = true; break;
      }
      p += 8 + len + (len & 1);
    }
    if (!have_data || fmt_tag != 1 || bits != 16) { delete s; return NULL; }
    info->samplerate = sr; info->channels = ch; info->format = SF_FORMAT_WAV | SF_FORMAT_PCM_16;
  }
  size_t n = data_len / 2; s->pcm.resize(n);
  for (size_t i = 0; i < n; i++) {
    unsigned lo = b[data_off + 2 * i], hi = b[data_off + 2 * i + 1];
    s->pcm[i] = (short)(big ? ((lo << 8) | hi) : ((hi << 8) | lo));
  }
  info->frames = (sf_count_t)(n / (info->channels > 0 ? info->channels : 1)); info->sections = 1; info->seekable = 1;
  return s;
}
static inline SNDFILE *sf_open(const char *path, int, SF_INFO *info) {
  FILE *fp = fopen(path, "rb"); if (!fp) return NULL;
  std::vector<unsigned char> b; unsigned char buf[65536]; size_t k;
  while ((k = fread(buf, 1, sizeof buf, fp)) > 0) b.insert(b.end(), buf, buf + k);
  fclose(fp);
  return shim_sf_from_bytes(b, info);
}
static inline SNDFILE *sf_open_fd(int fd, int, SF_INFO *info, int close_desc) {
  std::vector<unsigned char> b; unsigned char buf[65536]; ssize_t k;
  static std::vector<unsigned char> keep;  // a failed WAV probe must not lose the bytes for the RAW retry
  if (!keep.empty()) { b.swap(keep); }
  else while ((k = read(fd, buf, sizeof buf)) > 0) b.insert(b.end(), buf, buf + k);
  SNDFILE *s = shim_sf_from_bytes(b, info);
  if (!s) { keep.swap(b); return NULL; }
  s->fd = fd; s->close_fd = close_desc; return s;
}
static inline sf_count_t sf_read_short(SNDFILE *s, short *dst, sf_count_t n) {
  sf_count_t avail = (sf_count_t)s->pcm.size() - s->pos; if (avail < 0) avail = 0;
  if (n > avail) n = avail;
  if (n > 0) memcpy(dst, &s->pcm[s->pos], (size_t)n * sizeof(short));
  s->pos += n; return n;
}
static inline sf_count_t sf_seek(SNDFILE *s, sf_count_t frames, int whence) {
  if (whence != SEEK_SET || frames < 0) return -1;
  s->pos = frames; return frames;
}
static inline int sf_close(SNDFILE *s) { if (s->close_fd && s->fd >= 0) close(s->fd); delete s; return 0; }
#endif
