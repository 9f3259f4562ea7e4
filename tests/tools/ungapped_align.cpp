// TEST TOOL -- places short reads on a small reference without gaps, so that real reads with their real base qualities (the
// lambda read files of the reference's test suite) can go through the pileup path; breseq itself aligns with bowtie2, which is
// not in this image.  Not an aligner to be proud of: exact 11-mer seeds at three offsets, the whole read compared base by base
// (an N in the read matches anything), at most two mismatches, and a read is kept only if its best placement is the only one.
//
//   ungapped_align REFERENCE.fasta OUT.tsv READS.fastq.gz [READS.fastq.gz ...]
//   OUT.tsv: name <tab> flag (0 | 16) <tab> 0-based position <tab> bases as aligned <tab> qualities as aligned (Phred+33)
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

static const int K = 11, MAX_MISMATCHES = 2;

static bool pack(const char* s, uint32_t& key) {
  key = 0;
  for (int i = 0; i < K; ++i) {
    int v;
    switch (s[i]) { case 'A': v = 0; break; case 'C': v = 1; break; case 'G': v = 2; break; case 'T': v = 3; break; default: return false; }
    key = key << 2 | (uint32_t)v;
  }
  return true;
}

static bool next_line(gzFile f, std::string& line) {
  char buf[4096];
  if (!gzgets(f, buf, sizeof buf)) return false;
  line = buf;
  while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
  return true;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: ungapped_align REFERENCE.fasta OUT.tsv READS.fastq.gz ...\n"); return 2; }
  std::string ref;
  {
    FILE* f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 1; }
    char buf[4096];
    bool first = true;
    while (fgets(buf, sizeof buf, f)) {
      if (buf[0] == '>') { if (!first) break; first = false; continue; }
      for (char* p = buf; *p; ++p) if (*p > ' ') ref.push_back((char)toupper((unsigned char)*p));
    }
    fclose(f);
  }
  std::unordered_map<uint32_t, std::vector<int32_t>> index;
  for (size_t i = 0; i + K <= ref.size(); ++i) { uint32_t key; if (pack(ref.data() + i, key)) index[key].push_back((int32_t)i); }
  FILE* out = fopen(argv[2], "w");
  if (!out) { perror(argv[2]); return 1; }
  unsigned long kept = 0, seen = 0;
  for (int a = 3; a < argc; ++a) {
    gzFile f = gzopen(argv[a], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[a]); return 1; }
    std::string name, seq, plus, qual;
    while (next_line(f, name) && next_line(f, seq) && next_line(f, plus) && next_line(f, qual)) {
      ++seen;
      if (seq.size() != qual.size() || seq.size() < (size_t)K) continue;
      int best = MAX_MISMATCHES + 1, best_pos = -1, best_strand = 0, ties = 0;
      std::string oriented[2] = {seq, seq}, oriented_q[2] = {qual, qual};
      std::reverse(oriented[1].begin(), oriented[1].end());
      std::reverse(oriented_q[1].begin(), oriented_q[1].end());
      for (char& c : oriented[1]) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
      for (int strand = 0; strand < 2; ++strand) {
        const std::string& s = oriented[strand];
        for (size_t off : {(size_t)0, (s.size() - K) / 2, s.size() - K}) {
          uint32_t key;
          if (!pack(s.data() + off, key)) continue;
          auto hit = index.find(key);
          if (hit == index.end()) continue;
          for (int32_t at : hit->second) {
            const int64_t start = (int64_t)at - (int64_t)off;
            if (start < 0 || start + (int64_t)s.size() > (int64_t)ref.size()) continue;
            int mm = 0;
            for (size_t i = 0; i < s.size() && mm <= MAX_MISMATCHES; ++i) mm += s[i] != 'N' && s[i] != ref[(size_t)start + i];
            if (mm > MAX_MISMATCHES) continue;
            if (mm < best) { best = mm; best_pos = (int)start; best_strand = strand; ties = 1; }
            else if (mm == best && !((int)start == best_pos && strand == best_strand)) ++ties;
          }
        }
      }
      if (best_pos < 0 || ties != 1) continue;
      ++kept;
      fprintf(out, "%s\t%d\t%d\t%s\t%s\n", name.c_str() + 1, best_strand ? 16 : 0, best_pos, oriented[best_strand].c_str(), oriented_q[best_strand].c_str());
    }
    gzclose(f);
  }
  fclose(out);
  fprintf(stderr, "ungapped_align: %lu of %lu reads placed\n", kept, seen);
  return 0;
}
