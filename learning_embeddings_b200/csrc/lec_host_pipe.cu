// Host side of an end-to-end training iteration, issued from C (see include/lec_b200.h, lec_host_pipe_*).
//
// What the reference pays per iteration around the loss (order_embeddings_h.py:752-775, order_embeddings.py:619-635):
// the batch's indices come from host memory, the step runs, the scalar loss goes back to the host for logging.  Done
// from Python that is ~12 runtime calls per step (stream switch, copy, three event records / waits, the step, two
// copies back) and the interpreter between them: 100+ us of host time for a 65 us step, and eight ranks on one host
// contend for it.  Here ONE call enqueues the whole iteration:
//     copy stream:  [wait: slot's previous kernels done] -> H2D index block -> [Philox negative draw] -> event
//     main stream:  wait event -> lec_cone_step -> event (step done, slot free)
//     back stream:  wait (step done) -> D2H loss [+ error word] -> event (loss landed)
// `depth` slots rotate, so the copy of step i+1 overlaps the kernels of step i; the caller blocks (lec_host_pipe_wait)
// only when it wants a loss or needs a slot back.
#include <cuda_runtime.h>

#include <new>

#include "../../include/lec_b200.h"

namespace {
constexpr int kMaxDepth = 16;
}

struct lec_host_pipe {
    int depth;
    cudaStream_t copy, back;   // host->device side stream (+ the negative draw); device->host read-back stream
    cudaEvent_t copied[kMaxDepth], free_[kMaxDepth], loss[kMaxDepth];
    bool inflight[kMaxDepth];
};

extern "C" {

int lec_host_pipe_create(lec_host_pipe_t** out, int depth) {
    if (!out) return LEC_E_NULL;
    if (depth < 1 || depth > kMaxDepth) return LEC_E_SIZE;
    lec_host_pipe* p = new (std::nothrow) lec_host_pipe();
    if (!p) return LEC_E_SIZE;
    p->depth = depth;
    cudaError_t e = cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->back, cudaStreamNonBlocking);
    for (int i = 0; i < depth && e == cudaSuccess; ++i) {
        p->inflight[i] = false;
        e = cudaEventCreateWithFlags(&p->copied[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->free_[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->loss[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { delete p; return (int)e; }   // a failed create leaks at most a few events of a dying context
    *out = p;
    return 0;
}

void lec_host_pipe_destroy(lec_host_pipe_t* p) {
    if (!p) return;
    for (int i = 0; i < p->depth; ++i) {
        if (p->inflight[i]) cudaEventSynchronize(p->loss[i]);
        cudaEventDestroy(p->copied[i]);
        cudaEventDestroy(p->free_[i]);
        cudaEventDestroy(p->loss[i]);
    }
    cudaStreamDestroy(p->copy);
    cudaStreamDestroy(p->back);
    delete p;
}

int lec_host_pipe_submit(lec_host_pipe_t* p, int slot, const lec_step_t* step, const lec_host_sample_t* sample,
                         const void* host_block, int64_t bytes, void* dev_block, double* loss_host, int* err_host,
                         void* stream) {
    if (!p || !step || !loss_host) return LEC_E_NULL;
    if (slot < 0 || slot >= p->depth || bytes < 0) return LEC_E_SIZE;
    if (bytes > 0 && (!host_block || !dev_block)) return LEC_E_NULL;
    if (p->inflight[slot]) return LEC_E_SIZE;   // the caller has not collected this slot's previous step (lec_host_pipe_wait)
    cudaStream_t main = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    // the slot's staging area may still be read by the kernels of the step that used it last
    if (p->free_[slot]) e = cudaStreamWaitEvent(p->copy, p->free_[slot], 0);
    if (e == cudaSuccess && bytes > 0) e = cudaMemcpyAsync(dev_block, host_block, (size_t)bytes, cudaMemcpyHostToDevice, p->copy);
    if (e != cudaSuccess) return (int)e;
    if (sample) {
        // the draw needs the positives only -- not the table -- so it rides the copy stream and overlaps the kernels of
        // the step before
        if (!sample->graph) return LEC_E_NULL;
        const int rc = lec_sample_negatives_philox(sample->graph, step->pos_from, step->pos_to, step->idx_bytes, step->B, step->N,
                                                   sample->seed, sample->stream_id, const_cast<void*>(step->neg_to),
                                                   const_cast<void*>(step->neg_from), sample->status, (void*)p->copy);
        if (rc) return rc;
    }
    e = cudaEventRecord(p->copied[slot], p->copy);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(main, p->copied[slot], 0);
    if (e != cudaSuccess) return (int)e;
    if (const int rc = lec_cone_step(step, stream)) return rc;
    // The read-back rides its own stream behind the step's event: a device->host copy in the main stream would sit
    // between this step's update kernel and the next step's pair kernel (a few microseconds of copy-engine latency on
    // the critical path of every step).  The caller gives every slot its own *loss_step word (see lec_b200.h), so the
    // next steps do not overwrite it before it has been read.
    e = cudaEventRecord(p->free_[slot], main);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->back, p->free_[slot], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(loss_host, step->upd.loss_step, sizeof(double), cudaMemcpyDeviceToHost, p->back);
    if (e == cudaSuccess && err_host && step->xchg.world > 1 && step->xchg.error)
        e = cudaMemcpyAsync(err_host, step->xchg.error, sizeof(int), cudaMemcpyDeviceToHost, p->back);
    if (e == cudaSuccess) e = cudaEventRecord(p->loss[slot], p->back);
    if (e != cudaSuccess) return (int)e;
    p->inflight[slot] = true;
    return 0;
}

int lec_host_pipe_wait(lec_host_pipe_t* p, int slot) {
    if (!p) return LEC_E_NULL;
    if (slot < 0 || slot >= p->depth) return LEC_E_SIZE;
    if (!p->inflight[slot]) return 0;
    const cudaError_t e = cudaEventSynchronize(p->loss[slot]);
    p->inflight[slot] = false;
    return (int)e;
}

}  // extern "C"
