/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * Minimal implementation of the *legacy* pthreadpool API that the reference
 * QNNPACK calls from src/operator-run.c (pthreadpool_compute_{1d,1d_tiled,2d,
 * 2d_tiled,3d_tiled,4d_tiled}; declarations: torch/include/pthreadpool.h
 * "PTHREADPOOL_NO_DEPRECATED_API" block).  It exists only so that the compiled
 * reference in oracle/_ref/ can be timed on all host cores as the CPU baseline
 * (bench.py --impl reference) — the copies inside libtorch_cpu.so give correct
 * results but useless timings.
 *
 * Semantics: every tile start of the iteration space is visited
 * exactly once; the callback receives the tile start and the clipped tile size.
 * pool == NULL runs serially on the caller.  Work is handed out by an atomic
 * counter over the flattened tile index, so any number of threads is correct.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthreadpool.h>

struct job {
  int kind; /* 1:1d 2:1d_tiled 3:2d 4:2d_tiled 5:3d_tiled 6:4d_tiled */
  void* fn;
  void* ctx;
  size_t r[4];
  size_t t[4];
  size_t n[4]; /* tile counts per dim */
  size_t total;
};

struct pthreadpool {
  size_t nthreads; /* including the caller */
  pthread_t* workers;
  pthread_mutex_t mu;
  pthread_cond_t cv_start;
  pthread_cond_t cv_done;
  uint64_t epoch;
  int shutdown;
  size_t active;
  struct job job;
  atomic_size_t next;
};

static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }
static inline size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }

static void run_item(const struct job* j, size_t idx) {
  switch (j->kind) {
    case 1:
      ((pthreadpool_function_1d_t) j->fn)(j->ctx, idx);
      break;
    case 2: {
      const size_t i0 = idx * j->t[0];
      ((pthreadpool_function_1d_tiled_t) j->fn)(j->ctx, i0, min_sz(j->t[0], j->r[0] - i0));
      break;
    }
    case 3:
      ((pthreadpool_function_2d_t) j->fn)(j->ctx, idx / j->r[1], idx % j->r[1]);
      break;
    case 4: {
      const size_t i0 = (idx / j->n[1]) * j->t[0], j0 = (idx % j->n[1]) * j->t[1];
      ((pthreadpool_function_2d_tiled_t) j->fn)(
          j->ctx, i0, j0, min_sz(j->t[0], j->r[0] - i0), min_sz(j->t[1], j->r[1] - j0));
      break;
    }
    case 5: {
      const size_t k0 = (idx % j->n[2]) * j->t[2];
      const size_t rest = idx / j->n[2];
      const size_t j0 = (rest % j->n[1]) * j->t[1], i0 = (rest / j->n[1]) * j->t[0];
      ((pthreadpool_function_3d_tiled_t) j->fn)(
          j->ctx, i0, j0, k0, min_sz(j->t[0], j->r[0] - i0), min_sz(j->t[1], j->r[1] - j0),
          min_sz(j->t[2], j->r[2] - k0));
      break;
    }
    case 6: {
      const size_t l0 = (idx % j->n[3]) * j->t[3];
      size_t rest = idx / j->n[3];
      const size_t k0 = (rest % j->n[2]) * j->t[2];
      rest /= j->n[2];
      const size_t j0 = (rest % j->n[1]) * j->t[1], i0 = (rest / j->n[1]) * j->t[0];
      ((pthreadpool_function_4d_tiled_t) j->fn)(
          j->ctx, i0, j0, k0, l0, min_sz(j->t[0], j->r[0] - i0), min_sz(j->t[1], j->r[1] - j0),
          min_sz(j->t[2], j->r[2] - k0), min_sz(j->t[3], j->r[3] - l0));
      break;
    }
    default:
      break;
  }
}

static void drain(struct pthreadpool* p) {
  /* grab small batches to keep the atomic off the critical path */
  const size_t total = p->job.total;
  const size_t chunk = total / (p->nthreads * 16) + 1;
  for (;;) {
    const size_t b = atomic_fetch_add_explicit(&p->next, chunk, memory_order_relaxed);
    if (b >= total) break;
    const size_t e = min_sz(b + chunk, total);
    for (size_t i = b; i < e; i++) run_item(&p->job, i);
  }
}

static void* worker_main(void* arg) {
  struct pthreadpool* p = (struct pthreadpool*) arg;
  uint64_t seen = 0;
  for (;;) {
    pthread_mutex_lock(&p->mu);
    while (!p->shutdown && p->epoch == seen) pthread_cond_wait(&p->cv_start, &p->mu);
    if (p->shutdown) {
      pthread_mutex_unlock(&p->mu);
      return NULL;
    }
    seen = p->epoch;
    pthread_mutex_unlock(&p->mu);
    drain(p);
    pthread_mutex_lock(&p->mu);
    if (--p->active == 0) pthread_cond_signal(&p->cv_done);
    pthread_mutex_unlock(&p->mu);
  }
}

pthreadpool_t pthreadpool_create(size_t threads_count) {
  if (threads_count == 0) threads_count = 1;
  struct pthreadpool* p = (struct pthreadpool*) calloc(1, sizeof(*p));
  if (!p) return NULL;
  p->nthreads = threads_count;
  pthread_mutex_init(&p->mu, NULL);
  pthread_cond_init(&p->cv_start, NULL);
  pthread_cond_init(&p->cv_done, NULL);
  if (threads_count > 1) {
    p->workers = (pthread_t*) calloc(threads_count - 1, sizeof(pthread_t));
    for (size_t i = 0; i + 1 < threads_count; i++) pthread_create(&p->workers[i], NULL, worker_main, p);
  }
  return p;
}

size_t pthreadpool_get_threads_count(pthreadpool_t p) { return p ? p->nthreads : 1; }

void pthreadpool_destroy(pthreadpool_t p) {
  if (!p) return;
  pthread_mutex_lock(&p->mu);
  p->shutdown = 1;
  pthread_cond_broadcast(&p->cv_start);
  pthread_mutex_unlock(&p->mu);
  for (size_t i = 0; i + 1 < p->nthreads; i++) pthread_join(p->workers[i], NULL);
  free(p->workers);
  pthread_mutex_destroy(&p->mu);
  pthread_cond_destroy(&p->cv_start);
  pthread_cond_destroy(&p->cv_done);
  free(p);
}

static void dispatch(pthreadpool_t p, struct job* j) {
  if (j->total == 0) return;
  if (p == NULL || p->nthreads == 1) {
    for (size_t i = 0; i < j->total; i++) run_item(j, i);
    return;
  }
  pthread_mutex_lock(&p->mu);
  p->job = *j;
  atomic_store(&p->next, 0);
  p->active = p->nthreads - 1;
  p->epoch++;
  pthread_cond_broadcast(&p->cv_start);
  pthread_mutex_unlock(&p->mu);
  drain(p);
  pthread_mutex_lock(&p->mu);
  while (p->active != 0) pthread_cond_wait(&p->cv_done, &p->mu);
  pthread_mutex_unlock(&p->mu);
}

void pthreadpool_compute_1d(pthreadpool_t p, pthreadpool_function_1d_t fn, void* ctx, size_t range) {
  struct job j = {.kind = 1, .fn = (void*) fn, .ctx = ctx, .r = {range}, .total = range};
  dispatch(p, &j);
}

void pthreadpool_compute_1d_tiled(
    pthreadpool_t p, pthreadpool_function_1d_tiled_t fn, void* ctx, size_t range, size_t tile) {
  struct job j = {.kind = 2, .fn = (void*) fn, .ctx = ctx, .r = {range}, .t = {tile}};
  j.n[0] = ceil_div(range, tile);
  j.total = j.n[0];
  dispatch(p, &j);
}

void pthreadpool_compute_2d(
    pthreadpool_t p, pthreadpool_function_2d_t fn, void* ctx, size_t range_i, size_t range_j) {
  struct job j = {.kind = 3, .fn = (void*) fn, .ctx = ctx, .r = {range_i, range_j}};
  j.total = range_i * range_j;
  dispatch(p, &j);
}

void pthreadpool_compute_2d_tiled(
    pthreadpool_t p, pthreadpool_function_2d_tiled_t fn, void* ctx, size_t range_i, size_t range_j,
    size_t tile_i, size_t tile_j) {
  struct job j = {.kind = 4, .fn = (void*) fn, .ctx = ctx, .r = {range_i, range_j}, .t = {tile_i, tile_j}};
  j.n[0] = ceil_div(range_i, tile_i);
  j.n[1] = ceil_div(range_j, tile_j);
  j.total = j.n[0] * j.n[1];
  dispatch(p, &j);
}

void pthreadpool_compute_3d_tiled(
    pthreadpool_t p, pthreadpool_function_3d_tiled_t fn, void* ctx, size_t range_i, size_t range_j,
    size_t range_k, size_t tile_i, size_t tile_j, size_t tile_k) {
  struct job j = {
      .kind = 5, .fn = (void*) fn, .ctx = ctx, .r = {range_i, range_j, range_k}, .t = {tile_i, tile_j, tile_k}};
  j.n[0] = ceil_div(range_i, tile_i);
  j.n[1] = ceil_div(range_j, tile_j);
  j.n[2] = ceil_div(range_k, tile_k);
  j.total = j.n[0] * j.n[1] * j.n[2];
  dispatch(p, &j);
}

void pthreadpool_compute_4d_tiled(
    pthreadpool_t p, pthreadpool_function_4d_tiled_t fn, void* ctx, size_t range_i, size_t range_j,
    size_t range_k, size_t range_l, size_t tile_i, size_t tile_j, size_t tile_k, size_t tile_l) {
  struct job j = {
      .kind = 6,
      .fn = (void*) fn,
      .ctx = ctx,
      .r = {range_i, range_j, range_k, range_l},
      .t = {tile_i, tile_j, tile_k, tile_l}};
  j.n[0] = ceil_div(range_i, tile_i);
  j.n[1] = ceil_div(range_j, tile_j);
  j.n[2] = ceil_div(range_k, tile_k);
  j.n[3] = ceil_div(range_l, tile_l);
  j.total = j.n[0] * j.n[1] * j.n[2] * j.n[3];
  dispatch(p, &j);
}
