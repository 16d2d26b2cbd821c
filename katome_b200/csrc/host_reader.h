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
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <limits.h>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <system_error>
#include <thread>
#include <unistd.h>
#include <vector>

namespace ktg {

// One batch of raw sequences.  The bases live in pinned host memory when it can be had, so
// that ktg_add_reads copies them at PCIe speed and asynchronously.
struct ReadBatch {
    uint8_t *bases = nullptr;
    size_t size = 0, cap = 0;
    bool pinned = false;
    std::vector<uint64_t> offsets; // n_reads + 1
    ReadBatch() = default;
    ReadBatch(const ReadBatch &) = delete;
    ReadBatch &operator=(const ReadBatch &) = delete;
    ~ReadBatch() { release(); }
    uint64_t n_reads() const { return offsets.empty() ? 0 : offsets.size() - 1; }
    void release() {
        if (bases) {
            if (pinned) cudaFreeHost(bases);
            else free(bases);
        }
        bases = nullptr;
        size = cap = 0;
    }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        size_t ncap = cap ? cap : (1u << 20);
        while (ncap < want) ncap *= 2;
        uint8_t *nb = nullptr;
        bool np = cudaHostAlloc((void **)&nb, ncap, cudaHostAllocDefault) == cudaSuccess;
        if (!np) {
            (void)cudaGetLastError();
            nb = (uint8_t *)malloc(ncap);
            if (!nb) return false;
        }
        if (size) memcpy(nb, bases, size);
        const size_t keep = size;
        release();
        bases = nb;
        cap = ncap;
        size = keep;
        pinned = np;
        return true;
    }
    void clear() {
        size = 0;
        offsets.clear();
        offsets.push_back(0);
    }
    bool append(const char *s, size_t n) { // part of the current read
        if (size + n > cap && !reserve(size + n)) return false;
        memcpy(bases + size, s, n);
        size += n;
        return true;
    }
    void end_read() { offsets.push_back(size); }
};

// Reads a regular file in fixed blocks into page-locked buffers with several threads, ahead of its
// consumer: fread of a block from the page cache is a ~6 GB/s memcpy on one core, which made the host
// the pace setter of Build::create from a file (4 GB/s of FASTQ against 55 GB/s of PCIe behind it).
// Every buffer has HEAD bytes of room in front of the block, where the consumer puts the unfinished
// record of the previous block, so that a chunk is contiguous without moving the block.
class BlockReader {
  public:
    static constexpr size_t HEAD = 1u << 20;
    BlockReader(int fd, size_t file_size, size_t block, int n_buf, int n_threads)
        : fd_(fd), size_(file_size), block_(block), n_buf_(n_buf), n_threads_(n_threads) {}
    BlockReader(const BlockReader &) = delete;
    BlockReader &operator=(const BlockReader &) = delete;
    ~BlockReader() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
        for (uint8_t *p : buf_) cudaFreeHost(p);
    }
    bool init() {
        for (int i = 0; i < n_buf_; ++i) {
            uint8_t *p = nullptr;
            if (cudaHostAlloc((void **)&p, HEAD + block_ + 64, cudaHostAllocDefault) != cudaSuccess) {
                (void)cudaGetLastError();
                return false;
            }
            buf_.push_back(p);
            ready_.push_back(-1);
            free_for_.push_back(i);
            got_.push_back(0);
        }
        try {
            for (int t = 0; t < n_threads_; ++t) th_.emplace_back([this] { worker(); });
        } catch (const std::system_error &) {
            if (th_.empty()) return false;
        }
        return true;
    }
    size_t n_blocks() const { return (size_ + block_ - 1) / block_; }
    // block j, in order: its first byte (HEAD writable bytes in front of it) and its size
    bool acquire(size_t j, uint8_t **p, size_t *got) {
        const int s = (int)(j % n_buf_);
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return ready_[s] == (long long)j || failed_; });
        if (failed_) return false;
        *p = buf_[s] + HEAD;
        *got = got_[s];
        return true;
    }
    void release(size_t j) { // the buffer may be overwritten
        const int s = (int)(j % n_buf_);
        {
            std::lock_guard<std::mutex> lk(m_);
            free_for_[s] = (long long)j + n_buf_;
        }
        cv_.notify_all();
    }

  private:
    void worker() {
        for (;;) {
            const size_t j = next_.fetch_add(1);
            if (j >= n_blocks()) return;
            const int s = (int)(j % n_buf_);
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return free_for_[s] == (long long)j || stop_; });
                if (stop_) return;
            }
            size_t done = 0;
            const size_t want = std::min(block_, size_ - j * block_);
            bool bad = false;
            while (done < want) {
                const ssize_t r = pread(fd_, buf_[s] + HEAD + done, want - done, (off_t)(j * block_ + done));
                if (r < 0) {
                    bad = true;
                    break;
                }
                if (r == 0) break; // the file shrank: what is there is the block
                done += (size_t)r;
            }
            {
                std::lock_guard<std::mutex> lk(m_);
                got_[s] = done;
                ready_[s] = (long long)j;
                if (bad) failed_ = true;
            }
            cv_.notify_all();
        }
    }
    int fd_;
    size_t size_, block_;
    int n_buf_, n_threads_;
    std::vector<uint8_t *> buf_;
    std::vector<long long> ready_, free_for_;
    std::vector<size_t> got_;
    std::vector<std::thread> th_;
    std::atomic<size_t> next_{0};
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false, failed_ = false;
};

class ReadFile {
  public:
    ~ReadFile() {
        if (f_) fclose(f_);
        free(buf_);
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
        setvbuf(f_, nullptr, _IONBF, 0); // we do our own (large) buffering
        fasta_ = fasta;
        return true;
    }

    bool is_fasta() const { return fasta_; }
    // raw bytes for the device-side parser (must not be mixed with next_batch on one file)
    size_t read_raw(void *dst, size_t n) { return fread(dst, 1, n, f_); }
    int fd() const { return fileno(f_); }
    // size of a regular file, -1 for anything else (pipes, devices: those are read serially)
    long long regular_size() const {
        struct stat st;
        if (fstat(fileno(f_), &st) != 0 || !S_ISREG(st.st_mode)) return -1;
        return (long long)st.st_size;
    }

    // Appends records to a cleared batch until it holds >= max_bytes of bases.
    // Returns 1 if more records may follow, 0 at end of file, -1 on a malformed record.
    int next_batch(ReadBatch *out, size_t max_bytes, std::string *why) {
        out->clear();
        if (!out->reserve(max_bytes + (1u << 16))) {
            *why = "out of host memory";
            return -1;
        }
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

    // Next line (with its newline, like getline) as a view into the block buffer; false at end of
    // file.  Lines are found with memchr over 8 MiB blocks instead of one getline call per line.
    bool next_line(const char **p, size_t *n) {
        for (;;) {
            if (pos_ < end_) {
                const char *nl = (const char *)memchr(buf_ + pos_, '\n', end_ - pos_);
                if (nl) {
                    *p = buf_ + pos_;
                    *n = (size_t)(nl - (buf_ + pos_)) + 1;
                    pos_ += *n;
                    return true;
                }
                if (eof_) { // last line without a newline
                    *p = buf_ + pos_;
                    *n = end_ - pos_;
                    pos_ = end_;
                    return true;
                }
            }
            else if (eof_) return false;
            // keep the partial line, refill behind it
            const size_t keep = end_ - pos_;
            if (keep && pos_) memmove(buf_, buf_ + pos_, keep);
            pos_ = 0;
            end_ = keep;
            if (end_ + BLOCK > bcap_) {
                bcap_ = end_ + BLOCK;
                buf_ = (char *)realloc(buf_, bcap_);
            }
            const size_t got = fread(buf_ + end_, 1, BLOCK, f_);
            end_ += got;
            if (got < BLOCK) eof_ = true;
        }
    }

    int next_fastq(ReadBatch *out, size_t max_bytes, std::string *why) {
        const char *p;
        size_t n;
        while (out->size < max_bytes) {
            if (!next_line(&p, &n) || n == 0) return 0;
            if (p[0] != '@') {
                *why = "Expected @ at record start.";
                return -1;
            }
            if (next_line(&p, &n)) {
                if (!out->append(p, rtrim(p, n))) {
                    *why = "out of host memory";
                    return -1;
                }
            }
            out->end_read();
            next_line(&p, &n); // the '+' line is not looked at (rust-bio 0.10 reads and drops it)
            if (!next_line(&p, &n) || n == 0) {
                *why = "Incomplete record. Each FastQ record has to consist of 4 lines.";
                return -1;
            }
        }
        return 1;
    }

    int next_fasta(ReadBatch *out, size_t max_bytes, std::string *why) {
        if (!primed_) {
            have_ = next_line(&line_, &line_n_);
            primed_ = true;
        }
        while (out->size < max_bytes) {
            if (!have_ || line_n_ == 0) return 0;
            if (line_[0] != '>') {
                *why = "Expected > at record start.";
                return -1;
            }
            for (;;) {
                have_ = next_line(&line_, &line_n_);
                if (!have_ || line_n_ == 0 || line_[0] == '>') break;
                if (!out->append(line_, rtrim(line_, line_n_))) {
                    *why = "out of host memory";
                    return -1;
                }
            }
            out->end_read();
        }
        return 1;
    }

    static constexpr size_t BLOCK = 8u << 20;
    FILE *f_ = nullptr;
    bool fasta_ = false, primed_ = false, eof_ = false, have_ = false;
    char *buf_ = nullptr;
    size_t bcap_ = 0, pos_ = 0, end_ = 0;
    const char *line_ = nullptr;
    size_t line_n_ = 0;
};

} // namespace ktg
