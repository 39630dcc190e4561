// evx1_session.cpp -- encoder / decoder sessions behind the public API.  They mirror
// evx1_encoder_impl (evx1enc.cpp:11-168) and evx1_decoder_impl (evx1dec.cpp:10-135): stream
// header once, a 10-byte frame descriptor per frame, lazy initialisation on the first frame,
// frame counter, periodic intra.  Where the reference calls engine_encode_frame /
// engine_decode_frame (encode.cpp:205, decode.cpp:172) these sessions call the device
// library (include/evxgpu.h) for the pixel pipeline and entropy.cpp for the slice.
#include <stdlib.h>
#include <algorithm>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "entropy.h"
#include "evx1.h"
#include "evxgpu.h"

namespace evx {

namespace {

const uint16 kVersionWord = (2u << 8) | 47u;        // EVX_VERSION_WORD(2, 47), version.h:36-41

#pragma pack(push, 2)
struct stream_header          // evx_header, common.h:53-62 (14 bytes; byte 7 is padding)
{
    uint8 magic[4];
    uint16 size;
    uint8 ref_count;
    uint16 version;
    uint16 frame_width;
    uint16 frame_height;
};
struct frame_desc             // evx_frame, common.h:68-74 (10 bytes)
{
    uint32 type;              // EVX_FRAME_TYPE: 0 intra, 1 inter
    uint32 index;
    uint16 quality;
};
#pragma pack(pop)
static_assert(sizeof(stream_header) == 14, "evx_header is 14 bytes on the wire");
static_assert(sizeof(frame_desc) == 10, "evx_frame is 10 bytes on the wire");

inline int clip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

evx_status map_gpu_status(int rc)
{
    if (rc == 0) return EVX_SUCCESS;
    if (rc == 3) return EVX_ERROR_OUTOFMEMORY;
    if (rc == 1) return EVX_ERROR_INVALIDARG;
    return EVX_ERROR_HARDWAREFAIL;
}

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- encoder

class encoder_session : public evx1_encoder
{
    evx1_config cfg_;
    bool initialized_;
    frame_desc frame_;
    stream_header header_;
    evxgpu_handle *gpu_;
    slice_writer writer_;                          // table + records output: the stateful host binarisation (one frame at a time)
    evx1_frame_stats stats_;
    bool device_bins_, coder_threads_;

    // A frame between submit() and collect() is on the device (its kernels and copies are queued or running; the device
    // library holds two) or retired: a job that owns the frame's results and runs -- or has run -- the entropy stage.
    struct pending_frame
    {
        bool valid, first;
        frame_desc desc;
        double t_submit, gpu_ms, wait_ms;
        uint32 n_noncopy, d2h_bytes;
        uint64_t nbins;
    };
    enum { kDevMax = 16 };
    pending_frame dev_[kDevMax];                   // dev_[0] is the oldest of the frames on the device
    int dev_count_;                                // how many the device library takes: evxgpu_encode_capacity
    std::vector<evxgpu_block_desc> table_;         // retired frame, table + records output
    std::vector<int16> records_;

    // Bin-string output: the arithmetic coder of a slice needs nothing but the slice's bins (the coder is reset per
    // frame, serialize.cpp:323), so retired frames are coded by worker threads, several at a time, and collected in order.
    enum { kJobs = 12, kMaxWorkers = 8 };
    int kWorkers;                                  // evx1_config::coder_threads
    enum job_state { JOB_FREE = 0, JOB_QUEUED, JOB_RUNNING, JOB_DONE };
    struct job
    {
        job_state state;
        pending_frame f;
        std::vector<uint64_t> bins;
        slice_writer writer;                       // serialize_bins only: no state between frames
        uint32 bits;
        double ms;
    };
    job jobs_[kJobs];
    int head_, count_;                             // retired, uncollected frames: jobs head_, head_+1, ... (mod kJobs)
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::thread workers_[kMaxWorkers];
    bool threads_up_, stop_;

    void run_job(job &j)
    {
        const double t0 = now_ms();
        j.bits = device_bins_ ? j.writer.serialize_bins(j.bins.data(), j.f.nbins)
                              : writer_.serialize(table_.data(), records_.data(), j.f.n_noncopy);
        j.ms = now_ms() - t0;
    }

    void worker()
    {
        std::unique_lock<std::mutex> lk(m_);
        for (;;)
        {
            job *mine = NULL;
            for (int k = 0; k < count_ && !mine; ++k)
            {
                job &j = jobs_[(head_ + k) % kJobs];
                if (j.state == JOB_QUEUED) mine = &j;
            }
            if (!mine)
            {
                if (stop_) return;
                cv_work_.wait(lk);
                continue;
            }
            mine->state = JOB_RUNNING;
            lk.unlock();
            run_job(*mine);
            lk.lock();
            mine->state = JOB_DONE;
            cv_done_.notify_all();
        }
    }

    void start_threads()
    {
        if (threads_up_) return;
        stop_ = false;
        for (int k = 0; k < kWorkers; ++k) workers_[k] = std::thread(&encoder_session::worker, this);
        threads_up_ = true;
    }

    void stop_threads()
    {
        if (!threads_up_) return;
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_work_.notify_all();
        for (int k = 0; k < kWorkers; ++k) workers_[k].join();
        threads_up_ = false;
    }

    void drop_jobs()             // uncollected retired frames are forgotten; a slice being coded is waited for
    {
        std::unique_lock<std::mutex> lk(m_);
        for (int k = 0; k < count_; ++k)
        {
            job &j = jobs_[(head_ + k) % kJobs];
            if (j.state == JOB_QUEUED) j.state = JOB_FREE;
            while (j.state == JOB_RUNNING) cv_done_.wait(lk);
            j.state = JOB_FREE;
        }
        head_ = 0; count_ = 0;
    }

    int max_jobs() const { return device_bins_ ? kJobs : 1; }      // the host binarisation keeps one frame's table and records

    void clear_frame()          // clear_frame, common.cpp:50-64
    {
        frame_.type = 0;
        frame_.index = 0;
        frame_.quality = (uint16) clip(cfg_.default_quality, 1, 100);
    }

    evx_status initialize(uint32 width, uint32 height)      // evx1enc.cpp:66-90
    {
        if (initialized_) return EVX_ERROR_INVALID_RESOURCE;
        memset(&header_, 0, sizeof(header_));
        header_.magic[0] = 'E'; header_.magic[1] = 'V'; header_.magic[2] = 'X'; header_.magic[3] = '1';
        header_.ref_count = (uint8) cfg_.ref_count;
        header_.version = kVersionWord;
        header_.frame_width = (uint16) width;
        header_.frame_height = (uint16) height;
        header_.size = sizeof(stream_header);
        evxgpu_config gc = { cfg_.ref_count, cfg_.linear_quant, cfg_.deblocking, cfg_.frame_slots };
        int rc = evxgpu_create(cfg_.device, (int) width, (int) height, &gc, NULL, &gpu_);
        if (rc) return map_gpu_status(rc);
        // The device binarises the slice (include/evxgpu.h, evxgpu_set_output); EVX1_HOST_BINARISE=1 keeps
        // the table + records path and binarises here instead (same bits; for A/B measurements).
        device_bins_ = getenv("EVX1_HOST_BINARISE") == NULL;
        { const char *ct = getenv("EVX1_CODER_THREAD"); coder_threads_ = !(ct && ct[0] == '0'); }      // EVX1_CODER_THREAD=0: code on the caller's thread
        if (device_bins_ && (rc = evxgpu_set_output(gpu_, 1))) { evxgpu_destroy(gpu_); gpu_ = NULL; return map_gpu_status(rc); }
        int mbw = (int) ((width + 15) / 16), mbh = (int) ((height + 15) / 16);
        writer_.configure(mbw, mbh, cfg_.ref_count);
        for (int k = 0; k < kJobs; ++k) jobs_[k].writer.configure(mbw, mbh, cfg_.ref_count);
        table_.assign((size_t) mbw * mbh, evxgpu_block_desc());
        records_.assign((size_t) mbw * mbh * EVXGPU_MB_COEFFS, 0);
        initialized_ = true;
        return EVX_SUCCESS;
    }

    // Waits for the oldest frame on the device and moves its results into a job, which frees the device library's
    // slot.  async: the entropy stage starts on a worker thread; else it runs here.
    evx_status retire(bool async)
    {
        if (count_ >= max_jobs()) return EVX_ERROR_NOT_READY;
        pending_frame f = dev_[0];
        job &j = jobs_[(head_ + count_) % kJobs];
        int rc;
        const double tw = now_ms();
        if (device_bins_)
        {
            const uint64_t *bins = NULL;
            rc = evxgpu_encode_collect_bins(gpu_, &bins, &f.nbins, &f.n_noncopy);
            if (rc) return EVX_ERROR_EXECUTION_FAILURE;
            const size_t words = (size_t) ((f.nbins + 63) / 64);
            if (j.bins.size() < words) j.bins.resize(words + words / 2 + 64);
            memcpy(j.bins.data(), bins, words * 8);
        }
        else
        {
            rc = evxgpu_encode_collect(gpu_, table_.data(), records_.data(), &f.n_noncopy);
            if (rc) return EVX_ERROR_EXECUTION_FAILURE;
        }
        for (int k = 1; k < dev_count_; ++k) dev_[k - 1] = dev_[k];       // (only now: a failed collect keeps the frame)
        dev_count_--;
        f.d2h_bytes = (uint32) evxgpu_d2h_bytes(gpu_);
        f.gpu_ms = now_ms() - f.t_submit;
        f.wait_ms = now_ms() - tw;
        j.f = f; j.bits = 0; j.ms = 0.0;
        if (async && device_bins_ && coder_threads_)
        {
            start_threads();
            { std::lock_guard<std::mutex> g(m_); j.state = JOB_QUEUED; count_++; }
            cv_work_.notify_one();
        }
        else
        {
            run_job(j);
            std::lock_guard<std::mutex> g(m_);
            j.state = JOB_DONE; count_++;
        }
        return EVX_SUCCESS;
    }

public:
    explicit encoder_session(const evx1_config &cfg)
        : cfg_(cfg), initialized_(false), gpu_(NULL), device_bins_(true), coder_threads_(true), dev_count_(0), head_(0), count_(0), threads_up_(false), stop_(false)
    {
        kWorkers = cfg.coder_threads > 0 ? std::min<int>(cfg.coder_threads, kMaxWorkers) : 6;
        memset(&stats_, 0, sizeof(stats_));
        memset(&header_, 0, sizeof(header_));
        memset(dev_, 0, sizeof(dev_));
        for (int k = 0; k < kJobs; ++k) jobs_[k].state = JOB_FREE;
        clear_frame();
    }
    ~encoder_session() { clear(); stop_threads(); }

    evx_status clear()                                       // evx1enc.cpp:27-40
    {
        drop_jobs();
        if (!initialized_) return EVX_SUCCESS;
        clear_frame();
        if (gpu_) { evxgpu_destroy(gpu_); gpu_ = NULL; }      // drains the streams; uncollected frames are dropped
        dev_count_ = 0;
        initialized_ = false;
        return EVX_SUCCESS;
    }

    evx_status insert_intra() { frame_.type = 0; return EVX_SUCCESS; }                      // evx1enc.cpp:42-51

    evx_status set_quality(uint8 quality) { frame_.quality = (uint16) clip(quality, 1, 31); return EVX_SUCCESS; }   // evx1enc.cpp:53-64

    // First half of encode (evx1enc.cpp:92-156 up to and including the pixel part of engine_encode_frame,
    // encode.cpp:205-232): the frame goes to the device and the call returns.  The frame descriptor
    // (type, index, quality) is the session's state at this moment, exactly as encode() would have used it.
    evx_status submit(void *image, uint32 width, uint32 height)
    {
        if (!width || !height || !image) return EVX_ERROR_INVALIDARG;
        bool first = false;
        if (!initialized_)
        {
            if ((width & 1) || (height & 1) || width > 0xFFFF || height > 0xFFFF) return EVX_ERROR_EXECUTION_FAILURE;   // convert.cpp:126-130
            evx_status st = initialize(width, height);
            if (evx_failed(st)) return EVX_ERROR_EXECUTION_FAILURE;
            first = true;
        }
        if (width != header_.frame_width || height != header_.frame_height) return EVX_ERROR_INVALID_RESOURCE;
        const uint8 *rgb = static_cast<const uint8 *>(image);
        // With frames on the device the new frame's host->device copy starts right away, on the copy stream, under their
        // kernels -- before this thread waits for the oldest of them: a frame's kernels cannot start before its pixels are
        // there, and with three frames overlapping the new one is due the moment its slot is free.
        const bool early = dev_count_ >= 1 && !cfg_.device_frames;        // (a frame already on the device needs no upload)
        if (early && dev_count_ >= std::min<int>(kDevMax, evxgpu_encode_capacity(gpu_)) && count_ >= max_jobs()) return EVX_ERROR_NOT_READY;
        if (early && evxgpu_encode_upload(gpu_, rgb)) return EVX_ERROR_EXECUTION_FAILURE;
        while (dev_count_ > 0 && dev_count_ >= std::min<int>(kDevMax, evxgpu_encode_capacity(gpu_)))
        {   // the device holds all the frames it takes (two, or three overlapping ones; one with table + records output):
            // take the oldest one's results off it (it is finished or about to be); its arithmetic coder starts on a
            // worker thread while this thread queues the new frame
            if (count_ >= max_jobs()) return EVX_ERROR_NOT_READY;       // nowhere to retire it to: collect() first
            evx_status st = retire(true);
            if (evx_failed(st)) return st;
        }
        const double t0 = now_ms();
        int rc;
        if (early)
        {
            // the frame uploaded above; its kernels are queued right behind those of the frames still on the device
            rc = evxgpu_encode_submit(gpu_, NULL, 0, (int) frame_.type, frame_.index, (int) frame_.quality);
            while (rc == 8 && dev_count_ > 0)
            {
                evx_status st = retire(true);
                if (evx_failed(st)) return st;
                rc = evxgpu_encode_submit(gpu_, NULL, 0, (int) frame_.type, frame_.index, (int) frame_.quality);
            }
        }
        else
        {
            rc = evxgpu_encode_submit(gpu_, rgb, cfg_.device_frames ? 1 : 0, (int) frame_.type, frame_.index, (int) frame_.quality);
            while (rc == 8 && dev_count_ > 0)
            {
                evx_status st = retire(true);
                if (evx_failed(st)) return st;
                rc = evxgpu_encode_submit(gpu_, rgb, cfg_.device_frames ? 1 : 0, (int) frame_.type, frame_.index, (int) frame_.quality);
            }
        }
        if (rc) return EVX_ERROR_EXECUTION_FAILURE;
        pending_frame &f = dev_[dev_count_++];
        memset(&f, 0, sizeof(f));
        f.valid = true; f.first = first; f.desc = frame_; f.t_submit = t0;

        frame_.type = 1;                                                                   // EVX_ALLOW_INTER_FRAMES
        if (cfg_.periodic_intra > 0 && 0 == ((frame_.index + 1) % (uint32) cfg_.periodic_intra)) insert_intra();
        frame_.index++;
        return EVX_SUCCESS;
    }

    // Second half: stream header (first frame only), frame descriptor and the entropy-coded slice of the
    // oldest uncollected frame are appended to output -- the bytes encode() appends.
    evx_status collect(bit_stream *output)
    {
        if (!output) return EVX_ERROR_INVALIDARG;
        if (!count_)
        {
            if (!dev_count_) return EVX_ERROR_NOT_READY;
            evx_status st = retire(false);
            if (evx_failed(st)) return st;
        }
        job &j = jobs_[head_];
        {
            std::unique_lock<std::mutex> lk(m_);
            while (j.state != JOB_DONE) cv_done_.wait(lk);
        }
        const pending_frame f = j.f;
        const uint32 bits = j.bits;
        evx_status st = EVX_SUCCESS;
        if (!bits) st = EVX_ERROR_EXECUTION_FAILURE;
        if (evx_succeeded(st) && f.first && evx_failed(output->write_bytes(&header_, sizeof(header_)))) st = EVX_ERROR_EXECUTION_FAILURE;
        frame_desc desc = f.desc;
        if (evx_succeeded(st) && evx_failed(output->write_bytes(&desc, sizeof(desc)))) st = EVX_ERROR_EXECUTION_FAILURE;
        // serialize_slice's write failures are ignored by the reference; report ours
        if (evx_succeeded(st) && evx_failed(output->write_bits(const_cast<uint8 *>(device_bins_ ? j.writer.data() : writer_.data()), bits))) st = EVX_ERROR_EXECUTION_FAILURE;
        stats_.gpu_ms = f.gpu_ms; stats_.entropy_ms = j.ms; stats_.slice_bits = bits;
        stats_.noncopy_blocks = f.n_noncopy; stats_.d2h_bytes = f.d2h_bytes; stats_.wait_ms = f.wait_ms;
        {
            std::lock_guard<std::mutex> g(m_);
            j.state = JOB_FREE;
            head_ = (head_ + 1) % kJobs; count_--;
        }
        return st;
    }

    evx_status encode(void *image, uint32 width, uint32 height, bit_stream *output)        // evx1enc.cpp:92-156
    {
        if (!output || !width || !height || !image) return EVX_ERROR_INVALIDARG;
        if (dev_count_ || count_) return EVX_ERROR_NOT_READY;       // finish the pipelined frames with collect() first
        const frame_desc before = frame_;
        evx_status st = submit(image, width, height);
        if (evx_failed(st)) return st;
        st = collect(output);
        if (evx_failed(st)) frame_ = before;                 // the reference leaves its frame counter alone on failure
        return st;
    }

    // Debug views (evx1enc.cpp:170-305): an R8G8B8 picture of the source, the last reconstruction, or one of the
    // block-table visualisations.  Not available while a pipelined frame is uncollected (EVX_ERROR_NOT_READY).
    evx_status peek(EVX_PEEK_STATE peek_state, void *output)
    {
        if (!output) return EVX_ERROR_INVALIDARG;
        if (!initialized_) return EVX_SUCCESS;
        if (dev_count_ || count_) return EVX_ERROR_NOT_READY;
        const uint32 w = header_.frame_width, h = header_.frame_height, mbw = (w + 15) / 16;
        uint8 *out = static_cast<uint8 *>(output);
        switch (peek_state)
        {
            case EVX_PEEK_SOURCE:
                return evxgpu_peek_rgb(gpu_, 0, 0, out) ? EVX_ERROR_EXECUTION_FAILURE : EVX_SUCCESS;
            case EVX_PEEK_DESTINATION:
            {   // query_prediction_index_by_offset(frame, 1): the slot the last encoded frame was reconstructed into
                const uint32 R = (uint32) cfg_.ref_count;
                return evxgpu_peek_rgb(gpu_, 2, (int) ((frame_.index + R - 1) % R), out) ? EVX_ERROR_EXECUTION_FAILURE : EVX_SUCCESS;
            }
            case EVX_PEEK_BLOCK_TABLE: case EVX_PEEK_QUANT_TABLE: case EVX_PEEK_BLOCK_VARIANCE: case EVX_PEEK_SPMP_TABLE:
                break;
            default:
                return EVX_ERROR_NOTIMPL;                     // EVX_PEEK_PREDICTION has no case in the reference either
        }
        if (evxgpu_peek_table(gpu_, table_.data())) return EVX_ERROR_EXECUTION_FAILURE;
        for (uint32 j = 0; j < h; ++j)
        for (uint32 i = 0; i < w; ++i)
        {
            uint8 *px = out + ((size_t) j * w + i) * 3;
            const evxgpu_block_desc &e = table_[(size_t) (i / 16) + (size_t) (j / 16) * mbw];
            const bool intra = (e.block_type & 1) != 0, motion = (e.block_type & 2) != 0, copy = (e.block_type & 4) != 0;   // types.h:68-87
            if (peek_state == EVX_PEEK_BLOCK_TABLE) { px[0] = intra ? 255 : 0; px[1] = motion ? 255 : 0; px[2] = copy ? 255 : 0; }
            else if (peek_state == EVX_PEEK_SPMP_TABLE)
            {
                px[0] = 0;
                px[1] = e.sp_pred ? (uint8) (255 * e.sp_amount) : 0;
                px[2] = e.sp_pred ? (uint8) (255 * !e.sp_amount) : 0;
            }
            else if (copy) { px[0] = 255; px[1] = 0; px[2] = 0; }
            else if (peek_state == EVX_PEEK_QUANT_TABLE) px[0] = px[1] = px[2] = (uint8) (255 - 15 * e.q_index);
            else px[0] = px[1] = px[2] = (uint8) clip((int16) (e.variance / 30), 0, 255);
        }
        return EVX_SUCCESS;
    }

    evx_status last_frame_stats(evx1_frame_stats *out) { if (!out) return EVX_ERROR_INVALIDARG; *out = stats_; return EVX_SUCCESS; }
};

// ---------------------------------------------------------------- decoder

class decoder_session : public evx1_decoder
{
    evx1_config cfg_;
    bool initialized_;
    frame_desc frame_;
    stream_header header_;
    evxgpu_handle *gpu_;
    slice_reader reader_;
    std::vector<evxgpu_block_desc> table_;
    std::vector<int16> records_;
    evx1_frame_stats stats_;

    // A submitted frame is a job: its slice is entropy-decoded (slice_reader::parse, which needs nothing from other
    // frames) by a worker thread; collect() merges the oldest finished job into the stream's state (apply, in
    // frame order) and runs the pixel pipeline.  Up to kJobs frames may be uncollected.
    enum { kJobs = 12, kWorkers = 8 };     // a slice takes ≈2.4 ms to parse and ≈0.5 ms to apply and hand to the device: eight parsers keep the collecting thread busy
    enum job_state { JOB_FREE = 0, JOB_QUEUED, JOB_RUNNING, JOB_DONE };
    struct job
    {
        job_state state;
        std::vector<uint8> bytes;          // the slice (bit_stream bytes, absolute bit positions)
        uint32 pos, end;
        frame_desc desc;
        parsed_slice ps;
        int rc;
        double ms;
        bool on_device;                    // applied and submitted to the device (start()), its picture not yet collected
        uint32 n_noncopy;
        double apply_ms, t_start;
    };
    job jobs_[kJobs];
    int head_, count_;                     // uncollected jobs: head_, head_+1, ... (mod kJobs), in frame order
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::thread workers_[kWorkers];
    bool threads_up_, stop_;

    void clear_frame() { frame_.type = 0; frame_.index = 0; frame_.quality = (uint16) clip(cfg_.default_quality, 1, 100); }

    void run_job(job &j)
    {
        const double t0 = now_ms();
        j.rc = reader_.parse(j.bytes.data(), j.pos, j.end, j.ps);
        j.ms = now_ms() - t0;
    }

    void worker()
    {
        std::unique_lock<std::mutex> lk(m_);
        for (;;)
        {
            job *mine = NULL;
            for (int k = 0; k < count_ && !mine; ++k)
            {
                job &j = jobs_[(head_ + k) % kJobs];
                if (j.state == JOB_QUEUED) mine = &j;
            }
            if (!mine)
            {
                if (stop_) return;
                cv_work_.wait(lk);
                continue;
            }
            mine->state = JOB_RUNNING;
            lk.unlock();
            run_job(*mine);
            lk.lock();
            mine->state = JOB_DONE;
            cv_done_.notify_all();
        }
    }

    void start_threads()
    {
        if (threads_up_) return;
        stop_ = false;
        for (int k = 0; k < kWorkers; ++k) workers_[k] = std::thread(&decoder_session::worker, this);
        threads_up_ = true;
    }

    void stop_threads()
    {
        if (!threads_up_) return;
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_work_.notify_all();
        for (int k = 0; k < kWorkers; ++k) workers_[k].join();
        threads_up_ = false;
    }

    evx_status initialize(bit_stream *input)                 // evx1dec.cpp:41-69, verify_header common.cpp:25-43
    {
        if (initialized_) return EVX_ERROR_INVALID_RESOURCE;
        if (evx_failed(input->read_bytes(&header_, sizeof(header_)))) return EVX_ERROR_INVALID_RESOURCE;
        if (header_.magic[0] != 'E' || header_.magic[1] != 'V' || header_.magic[2] != 'X' || header_.magic[3] != '1') return EVX_ERROR_INVALID_RESOURCE;
        if (header_.version != kVersionWord || header_.size != sizeof(stream_header)) return EVX_ERROR_INVALID_RESOURCE;
        // the reference rejects any ring size but its compiled one; here the header decides
        if (header_.ref_count < 2 || header_.ref_count > 8) return EVX_ERROR_INVALID_RESOURCE;
        if (!header_.frame_width || !header_.frame_height || (header_.frame_width & 1) || (header_.frame_height & 1)) return EVX_ERROR_INVALID_RESOURCE;
        evxgpu_config gc = { header_.ref_count, cfg_.linear_quant, cfg_.deblocking, 0 };
        int rc = evxgpu_create(cfg_.device, header_.frame_width, header_.frame_height, &gc, NULL, &gpu_);
        if (rc) return map_gpu_status(rc);
        int mbw = (header_.frame_width + 15) / 16, mbh = (header_.frame_height + 15) / 16;
        reader_.configure(mbw, mbh, header_.ref_count);
        evxgpu_block_desc zero;
        memset(&zero, 0, sizeof(zero));
        table_.assign((size_t) mbw * mbh, zero);             // aligned_zero_memory, common.cpp:146
        records_.assign((size_t) mbw * mbh * EVXGPU_MB_COEFFS, 0);
        initialized_ = true;
        return EVX_SUCCESS;
    }

    // header (first frame), frame descriptor and the slice's bytes into a job; the caller decides who parses it
    evx_status enqueue(bit_stream *input, job **out)
    {
        if (count_ >= kJobs) return EVX_ERROR_NOT_READY;
        if (!initialized_ && evx_failed(initialize(input))) return EVX_ERROR_EXECUTION_FAILURE;
        frame_desc incoming;
        if (evx_failed(input->read_bytes(&incoming, sizeof(incoming)))) return EVX_ERROR_EXECUTION_FAILURE;
        if (incoming.index != frame_.index) return EVX_ERROR_EXECUTION_FAILURE;          // evx1dec.cpp:77-80
        frame_ = incoming;
        job &j = jobs_[(head_ + count_) % kJobs];
        j.pos = input->query_read_index(); j.end = input->query_write_index();
        const size_t nbytes = ((size_t) j.end + 7) >> 3;
        if (j.bytes.size() < nbytes) j.bytes.resize(nbytes + nbytes / 2 + 64);
        memcpy(j.bytes.data(), input->query_data(), nbytes);
        j.desc = frame_; j.rc = 0; j.ms = 0.0; j.on_device = false; j.n_noncopy = 0; j.apply_ms = 0.0; j.t_start = 0.0;
        frame_.index++;
        input->empty();                                                                    // evx1dec.cpp:120
        *out = &j;
        return EVX_SUCCESS;
    }

public:
    explicit decoder_session(const evx1_config &cfg) : cfg_(cfg), initialized_(false), gpu_(NULL), head_(0), count_(0), threads_up_(false), stop_(false)
    {
        memset(&header_, 0, sizeof(header_));
        memset(&stats_, 0, sizeof(stats_));
        for (int k = 0; k < kJobs; ++k) jobs_[k].state = JOB_FREE;
        clear_frame();
    }
    ~decoder_session() { clear(); stop_threads(); }

    evx_status clear()                                       // evx1dec.cpp:27-39
    {
        {   // uncollected frames are dropped; a slice being parsed is waited for
            std::unique_lock<std::mutex> lk(m_);
            for (int k = 0; k < count_; ++k)
            {
                job &j = jobs_[(head_ + k) % kJobs];
                if (j.state == JOB_QUEUED) j.state = JOB_FREE;
                while (j.state == JOB_RUNNING) cv_done_.wait(lk);
                j.state = JOB_FREE;
            }
            head_ = 0; count_ = 0;
        }
        if (!initialized_) return EVX_SUCCESS;
        clear_frame();
        if (gpu_) { evxgpu_destroy(gpu_); gpu_ = NULL; }
        initialized_ = false;
        return EVX_SUCCESS;
    }

    // First half of decode (evx1dec.cpp:87-123 up to unserialize_slice, decode.cpp:172-180): the frame is taken
    // out of `input` (which is emptied, as decode() does) and handed to a parser thread; the call returns.
    evx_status submit(bit_stream *input)
    {
        if (!input) return EVX_ERROR_INVALIDARG;
        job *j = NULL;
        evx_status st = enqueue(input, &j);
        if (evx_failed(st)) return st;
        start_threads();
        {
            std::lock_guard<std::mutex> g(m_);
            j->state = JOB_QUEUED;
            count_++;
        }
        cv_work_.notify_one();
        return EVX_SUCCESS;
    }

    // Second half (the rest of unserialize_slice's effect, decode_slice, deblocking, colour conversion:
    // decode.cpp:182-198): the oldest submitted frame's RGB8 picture into output.
    evx_status collect(void *output)
    {
        if (!output) return EVX_ERROR_INVALIDARG;
        job *j, *next = NULL;
        {
            std::unique_lock<std::mutex> lk(m_);
            if (!count_) return EVX_ERROR_NOT_READY;
            j = &jobs_[head_];
            while (j->state != JOB_DONE) cv_done_.wait(lk);
            if (count_ >= 2 && jobs_[(head_ + 1) % kJobs].state == JOB_DONE) next = &jobs_[(head_ + 1) % kJobs];
        }
        // The oldest frame's picture leaves the device on the copy stream; under that copy the next parsed frame is merged
        // into the stream's state and handed to the device (apply is in frame order: this frame's came first).
        evx_status st = j->on_device ? EVX_SUCCESS : start(*j);
        if (evx_succeeded(st) && evxgpu_decode_collect_begin(gpu_, static_cast<uint8 *>(output), cfg_.device_frames ? 1 : 0)) st = EVX_ERROR_EXECUTION_FAILURE;
        if (evx_succeeded(st) && next && !next->on_device && evx_failed(start(*next))) next->rc = 1;      // reported by that frame's collect
        if (evx_succeeded(st) && evxgpu_decode_collect_end(gpu_)) st = EVX_ERROR_EXECUTION_FAILURE;
        if (evx_succeeded(st)) finish_stats(*j);
        {
            std::lock_guard<std::mutex> g(m_);
            j->state = JOB_FREE;
            head_ = (head_ + 1) % kJobs; count_--;
        }
        return st;
    }

    evx_status decode(bit_stream *input, void *output)       // evx1dec.cpp:87-123
    {
        if (!input || !output) return EVX_ERROR_INVALIDARG;
        if (count_) return EVX_ERROR_NOT_READY;               // finish the pipelined frames with collect() first
        job *j = NULL;
        evx_status st = enqueue(input, &j);
        if (evx_failed(st)) return st;
        run_job(*j);                                          // one frame at a time: parsed right here
        st = start(*j);
        if (evx_failed(st)) return st;
        if (evxgpu_decode_collect(gpu_, static_cast<uint8 *>(output), cfg_.device_frames ? 1 : 0)) return EVX_ERROR_EXECUTION_FAILURE;
        finish_stats(*j);
        return EVX_SUCCESS;
    }

    evx_status last_frame_stats(evx1_frame_stats *out) { if (!out) return EVX_ERROR_INVALIDARG; *out = stats_; return EVX_SUCCESS; }

private:
    // unserialize_slice's in-order half (apply) and the frame's kernels queued on the device
    evx_status start(job &j)
    {
        if (j.rc) return EVX_ERROR_EXECUTION_FAILURE;
        uint32 n_noncopy = 0;
        const double t0 = now_ms();
        int16 *rec = j.ps.records.empty() ? records_.data() : j.ps.records.data();          // resolved in place: no copy of the coefficients
        if (reader_.apply(j.ps, table_.data(), rec, &n_noncopy)) return EVX_ERROR_EXECUTION_FAILURE;
        const double t1 = now_ms();
        if (evxgpu_decode_submit(gpu_, table_.data(), rec, n_noncopy, (int) j.desc.type, j.desc.index)) return EVX_ERROR_EXECUTION_FAILURE;
        j.on_device = true; j.n_noncopy = n_noncopy; j.apply_ms = t1 - t0; j.t_start = t1;
        return EVX_SUCCESS;
    }

    void finish_stats(const job &j)
    {
        stats_.entropy_ms = j.ms + j.apply_ms; stats_.gpu_ms = now_ms() - j.t_start; stats_.noncopy_blocks = j.n_noncopy;
        stats_.slice_bits = j.end - j.pos; stats_.d2h_bytes = (uint32) ((size_t) header_.frame_width * header_.frame_height * 3);
        stats_.wait_ms = 0.0;
    }
};

}  // namespace

void default_config(evx1_config *cfg)                        // config.h:38-53
{
    cfg->device = 0;
    cfg->ref_count = 4;
    cfg->linear_quant = 0;
    cfg->deblocking = 1;
    cfg->periodic_intra = 3600;
    cfg->default_quality = 8;
    cfg->frame_slots = 0;
    cfg->coder_threads = 0;
    cfg->device_frames = 0;
}

static bool config_ok(const evx1_config &c) { return c.ref_count >= 2 && c.ref_count <= 8 && c.device >= 0; }

evx_status create_encoder_ex(const evx1_config &cfg, evx1_encoder **output)
{
    if (!output || !config_ok(cfg)) return EVX_ERROR_INVALIDARG;
    *output = new (std::nothrow) encoder_session(cfg);
    return *output ? EVX_SUCCESS : EVX_ERROR_OUTOFMEMORY;
}

evx_status create_decoder_ex(const evx1_config &cfg, evx1_decoder **output)
{
    if (!output || !config_ok(cfg)) return EVX_ERROR_INVALIDARG;
    *output = new (std::nothrow) decoder_session(cfg);
    return *output ? EVX_SUCCESS : EVX_ERROR_OUTOFMEMORY;
}

evx_status create_encoder(evx1_encoder **output) { evx1_config c; default_config(&c); return create_encoder_ex(c, output); }   // evx1.cpp:8-24
evx_status create_decoder(evx1_decoder **output) { evx1_config c; default_config(&c); return create_decoder_ex(c, output); }   // evx1.cpp:26-42

evx_status destroy_encoder(evx1_encoder *input)              // evx1.cpp:44-60
{
    if (!input) return EVX_ERROR_INVALIDARG;
    delete static_cast<encoder_session *>(input);
    return EVX_SUCCESS;
}

evx_status destroy_decoder(evx1_decoder *input)              // evx1.cpp:62-78
{
    if (!input) return EVX_ERROR_INVALIDARG;
    delete static_cast<decoder_session *>(input);
    return EVX_SUCCESS;
}

}  // namespace evx
