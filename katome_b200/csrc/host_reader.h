// host_reader.h -- the reads batcher: FASTQ / FASTA record reader feeding
// batches of raw sequences to ktg_add_reads.
//
// Replaces check_files (algorithms/builder.rs:57-77) and the record iteration
// of create_fastq / create_fasta (builder.rs:118-165), whose parsing is done by
// the un-vendored crate rust-bio 0.10.0 (Cargo.lock:22-25).  Behaviour kept:
//   FASTQ: strict 4-line records; a header that does not start with '@' is an
//          error; seq() is line 2 with trailing whitespace trimmed; an empty
//          4th line is an "incomplete record" error; sequence and quality
//          lengths are not compared; end of file at a record boundary ends
//          the iteration.
//   FASTA: header must start with '>'; sequence lines are concatenated after
//          trimming trailing whitespace.
// The ACGT filter, byte total and length check are NOT done here: they run on
// the GPU (pack_reads_kernel), exactly once, for every input path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits.h>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace ktg {

struct ReadBatch {
    std::vector<uint8_t> bases;
    std::vector<uint64_t> offsets; // n_reads + 1
    uint64_t n_reads() const { return offsets.empty() ? 0 : offsets.size() - 1; }
    void clear() {
        bases.clear();
        offsets.clear();
        offsets.push_back(0);
    }
    void push(const char *s, size_t n) {
        bases.insert(bases.end(), (const uint8_t *)s, (const uint8_t *)s + n);
        offsets.push_back(bases.size());
    }
};

class ReadFile {
  public:
    ~ReadFile() {
        if (f_) fclose(f_);
        free(line_);
        free(aux_);
    }

    // check_files: canonicalize, reject directories and missing files
    bool open(const char *path, bool fasta, std::string *why) {
        char real[PATH_MAX];
        if (!realpath(path, real)) {
            *why = std::string("Coulndt resolve path: ") + path; // sic, builder.rs:62
            return false;
        }
        struct stat st;
        if (stat(real, &st) != 0) {
            *why = std::string(real) + " does not exist";
            return false;
        }
        if (S_ISDIR(st.st_mode)) {
            *why = std::string(real) + " is a directory";
            return false;
        }
        f_ = fopen(real, "rb");
        if (!f_) {
            *why = std::string("Couldn't open all files: ") + real;
            return false;
        }
        fasta_ = fasta;
        return true;
    }

    // Appends records to a cleared batch until it holds >= max_bytes of bases.
    // Returns 1 if more records may follow, 0 at end of file, -1 on a malformed record.
    int next_batch(ReadBatch *out, size_t max_bytes, std::string *why) {
        out->clear();
        return fasta_ ? next_fasta(out, max_bytes, why) : next_fastq(out, max_bytes, why);
    }

  private:
    static size_t rtrim(const char *s, size_t n) {
        while (n > 0) {
            char c = s[n - 1];
            if (c == '\n' || c == '\r' || c == ' ' || c == '\t' || c == '\v' || c == '\f') --n;
            else break;
        }
        return n;
    }

    int next_fastq(ReadBatch *out, size_t max_bytes, std::string *why) {
        while (out->bases.size() < max_bytes) {
            ssize_t lh = getline(&line_, &cap_, f_);
            if (lh <= 0) return 0;
            if (line_[0] != '@') {
                *why = "Expected @ at record start.";
                return -1;
            }
            ssize_t ls = getline(&line_, &cap_, f_);
            if (ls < 0) ls = 0;
            size_t len = rtrim(line_, (size_t)ls);
            out->push(line_, len);
            ssize_t lp = getline(&aux_, &aux_cap_, f_);
            (void)lp;
            ssize_t lq = getline(&aux_, &aux_cap_, f_);
            if (lq <= 0) {
                *why = "Incomplete record. Each FastQ record has to consist of 4 lines.";
                return -1;
            }
        }
        return 1;
    }

    int next_fasta(ReadBatch *out, size_t max_bytes, std::string *why) {
        if (!primed_) {
            have_ = getline(&line_, &cap_, f_);
            primed_ = true;
        }
        while (out->bases.size() < max_bytes) {
            if (have_ <= 0) return 0;
            if (line_[0] != '>') {
                *why = "Expected > at record start.";
                return -1;
            }
            size_t start = out->bases.size();
            for (;;) {
                have_ = getline(&line_, &cap_, f_);
                if (have_ <= 0 || line_[0] == '>') break;
                size_t m = rtrim(line_, (size_t)have_);
                out->bases.insert(out->bases.end(), (const uint8_t *)line_, (const uint8_t *)line_ + m);
            }
            (void)start;
            out->offsets.push_back(out->bases.size());
        }
        return 1;
    }

    FILE *f_ = nullptr;
    bool fasta_ = false, primed_ = false;
    char *line_ = nullptr, *aux_ = nullptr;
    size_t cap_ = 0, aux_cap_ = 0;
    ssize_t have_ = 0;
};

} // namespace ktg
