/*
 * mpi.h -- single-node stand-in for the handful of MPI entry points that the
 * phyNGSC driver and our own host driver use.  It exists because the build
 * image ships no MPI (no mpicxx / mpiexec / libmpi): ranks are fork()ed
 * children of the launching process, collectives go through one anonymous
 * MAP_SHARED arena, and MPI-IO maps onto pread/pwrite.
 *
 *   PHY_SHIM_NP=<n> ./program args...     # plays the role of `mpiexec -np n`
 *
 * With a real MPI installed, compile against its <mpi.h> instead (drop the -I
 * of this directory); nothing in the callers depends on the shim.
 *
 * Header-only, C and C++.  Not a general MPI: one communicator, blocking
 * collectives only, datatypes identified by their byte size.
 */
#ifndef PHY_MPI_SHIM_H
#define PHY_MPI_SHIM_H

#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype; /* handle == extent in bytes */
typedef int MPI_Info;
typedef int MPI_Op;
typedef long long MPI_Offset;
typedef long MPI_Aint;
typedef struct { int count_bytes; } MPI_Status;
typedef struct phy_shim_file { int fd; int slot; } *MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_SUCCESS 0
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_MODE_RDONLY 1
#define MPI_MODE_RDWR 2
#define MPI_MODE_WRONLY 4
#define MPI_MODE_CREATE 8
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_DOUBLE 8
#define MPI_LONG_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_SUM 1
#define MPI_MAX 2

#define PHY_SHIM_MAX_RANKS 64
#define PHY_SHIM_MAX_FILES 8
#define PHY_SHIM_SLOT_BYTES (1u << 20)

struct phy_shim_arena {
  pthread_barrier_t barrier;
  pthread_mutex_t lock;
  long long shared_off[PHY_SHIM_MAX_FILES];
  volatile int abort_code;                 /* set by MPI_Abort before it signals the other ranks */
  volatile pid_t pids[PHY_SHIM_MAX_RANKS]; /* every rank's process id (written by the rank itself) */
  unsigned char slots[PHY_SHIM_MAX_RANKS][PHY_SHIM_SLOT_BYTES];
};

static struct phy_shim_arena *phy_shim_arena_p = 0;
static int phy_shim_rank = 0, phy_shim_np = 1, phy_shim_files_open = 0;
static pid_t phy_shim_kids[PHY_SHIM_MAX_RANKS];

/* MPI_Abort ends every rank: the aborting rank signals the others, which leave with the same exit code */
static void phy_shim_on_term(int sig) { (void)sig; _exit(phy_shim_arena_p && phy_shim_arena_p->abort_code ? phy_shim_arena_p->abort_code : 143); }

static inline int MPI_Init_thread(int *argc, char ***argv, int required, int *provided) {
  (void)argc; (void)argv;
  const char *e = getenv("PHY_SHIM_NP");
  phy_shim_np = e ? atoi(e) : 1;
  if (phy_shim_np < 1 || phy_shim_np > PHY_SHIM_MAX_RANKS) {
    fprintf(stderr, "mpi shim: PHY_SHIM_NP must be in 1..%d\n", PHY_SHIM_MAX_RANKS);
    exit(97);
  }
  phy_shim_arena_p = (struct phy_shim_arena *)mmap(0, sizeof(struct phy_shim_arena), PROT_READ | PROT_WRITE,
                                                   MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (phy_shim_arena_p == MAP_FAILED) { perror("mpi shim: mmap"); exit(97); }
  pthread_barrierattr_t ba;
  pthread_barrierattr_init(&ba);
  pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&phy_shim_arena_p->barrier, &ba, (unsigned)phy_shim_np);
  pthread_mutexattr_t ma;
  pthread_mutexattr_init(&ma);
  pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
  pthread_mutex_init(&phy_shim_arena_p->lock, &ma);
  fflush(stdout); fflush(stderr);
  for (int r = 1; r < phy_shim_np; ++r) {
    pid_t k = fork();
    if (k < 0) { perror("mpi shim: fork"); exit(97); }
    if (k == 0) { phy_shim_rank = r; break; }
    phy_shim_kids[r] = k;
  }
  phy_shim_arena_p->pids[phy_shim_rank] = getpid();
  signal(SIGTERM, phy_shim_on_term);
  if (provided) *provided = required;
  return MPI_SUCCESS;
}
static inline int MPI_Init(int *argc, char ***argv) { int p; return MPI_Init_thread(argc, argv, 0, &p); }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = phy_shim_rank; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; *n = phy_shim_np; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; pthread_barrier_wait(&phy_shim_arena_p->barrier); return 0; }
static inline double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Finalize(void) {
  fflush(stdout); fflush(stderr);
  if (phy_shim_rank == 0) {
    int worst = 0;
    for (int r = 1; r < phy_shim_np; ++r) {
      int st = 0;
      if (waitpid(phy_shim_kids[r], &st, 0) > 0 && !(WIFEXITED(st) && WEXITSTATUS(st) == 0)) worst = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
    }
    if (worst) _exit(worst); /* a rank that failed after its last collective still fails the run */
  }
  return 0;
}
static inline int MPI_Abort(MPI_Comm c, int code) {
  (void)c; fflush(stdout); fflush(stderr);
  if (phy_shim_arena_p) {
    phy_shim_arena_p->abort_code = code ? code : 1;
    for (int r = 0; r < phy_shim_np; ++r) { const pid_t p = phy_shim_arena_p->pids[r]; if (r != phy_shim_rank && p > 0) kill(p, SIGTERM); }
  }
  _exit(code);
  return 0;
}

/* ---- MPI-IO ------------------------------------------------------------ */
static inline int MPI_File_open(MPI_Comm c, const char *name, int amode, MPI_Info info, MPI_File *fh) {
  (void)c; (void)info;
  int fl = 0;
  if (amode & MPI_MODE_RDWR) fl |= O_RDWR; else if (amode & MPI_MODE_WRONLY) fl |= O_WRONLY; else fl |= O_RDONLY;
  if (amode & MPI_MODE_CREATE) fl |= O_CREAT;
  int slot = phy_shim_files_open++;
  if (slot >= PHY_SHIM_MAX_FILES) return 1;
  /* collective: rank 0 creates, the others open afterwards */
  int fd = -1;
  if (phy_shim_rank == 0) { fd = open(name, fl, 0644); phy_shim_arena_p->shared_off[slot] = 0; }
  MPI_Barrier(0);
  if (phy_shim_rank != 0) fd = open(name, fl & ~O_CREAT, 0644);
  if (fd < 0) { *fh = 0; return 1; }
  *fh = (MPI_File)malloc(sizeof(struct phy_shim_file));
  (*fh)->fd = fd; (*fh)->slot = slot;
  return 0;
}
static inline int MPI_File_close(MPI_File *fh) { if (*fh) { close((*fh)->fd); free(*fh); *fh = 0; } return 0; }
static inline int MPI_File_get_size(MPI_File fh, MPI_Offset *sz) {
  struct stat st; if (fstat(fh->fd, &st)) return 1; *sz = (MPI_Offset)st.st_size; return 0;
}
static inline int phy_shim_prw(int wr, int fd, void *buf, long long n, long long off) {
  char *p = (char *)buf; long long done = 0;
  while (done < n) {
    ssize_t k = wr ? pwrite(fd, p + done, (size_t)(n - done), (off_t)(off + done))
                   : pread(fd, p + done, (size_t)(n - done), (off_t)(off + done));
    if (k <= 0) break; /* short read at EOF leaves the tail untouched, like MPI */
    done += k;
  }
  return 0;
}
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void *buf, long long count, MPI_Datatype t, MPI_Status *s) {
  (void)s; return phy_shim_prw(0, fh->fd, buf, count * (long long)t, off);
}
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void *buf, long long count, MPI_Datatype t, MPI_Status *s) {
  (void)s; return phy_shim_prw(1, fh->fd, (void *)buf, count * (long long)t, off);
}
static inline int MPI_File_write_shared(MPI_File fh, const void *buf, long long count, MPI_Datatype t, MPI_Status *s) {
  (void)s;
  long long n = count * (long long)t, off;
  pthread_mutex_lock(&phy_shim_arena_p->lock);
  off = phy_shim_arena_p->shared_off[fh->slot];
  phy_shim_arena_p->shared_off[fh->slot] = off + n;
  pthread_mutex_unlock(&phy_shim_arena_p->lock);
  return phy_shim_prw(1, fh->fd, (void *)buf, n, off);
}

/* ---- derived datatypes (extent only) ----------------------------------- */
static inline int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)(intptr_t)p; return 0; }
static inline int MPI_Type_create_struct(int n, const int *bl, const MPI_Aint *disp, const MPI_Datatype *ty, MPI_Datatype *out) {
  long ext = 0;
  for (int i = 0; i < n; ++i) { long e = (long)disp[i] + (long)bl[i] * ty[i]; if (e > ext) ext = e; }
  *out = (int)ext; return 0;
}
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return 0; }

/* ---- collectives -------------------------------------------------------- */
static inline void phy_shim_check(long long bytes) {
  if (bytes > (long long)PHY_SHIM_SLOT_BYTES) { fprintf(stderr, "mpi shim: message of %lld B exceeds slot\n", bytes); _exit(97); }
}
static inline int MPI_Gatherv(const void *sb, int sc, MPI_Datatype st, void *rb, const int *rc, const int *displs,
                              MPI_Datatype rt, int root, MPI_Comm c) {
  (void)c;
  long long n = (long long)sc * st; phy_shim_check(n);
  memcpy(phy_shim_arena_p->slots[phy_shim_rank], sb, (size_t)n);
  MPI_Barrier(0);
  if (phy_shim_rank == root)
    for (int r = 0; r < phy_shim_np; ++r)
      memcpy((char *)rb + (long long)displs[r] * rt, phy_shim_arena_p->slots[r], (size_t)((long long)rc[r] * rt));
  MPI_Barrier(0);
  return 0;
}
static inline int MPI_Gather(const void *sb, int sc, MPI_Datatype st, void *rb, int rc, MPI_Datatype rt, int root, MPI_Comm c) {
  (void)c;
  long long n = (long long)sc * st; phy_shim_check(n);
  memcpy(phy_shim_arena_p->slots[phy_shim_rank], sb, (size_t)n);
  MPI_Barrier(0);
  if (phy_shim_rank == root)
    for (int r = 0; r < phy_shim_np; ++r) memcpy((char *)rb + (long long)r * rc * rt, phy_shim_arena_p->slots[r], (size_t)((long long)rc * rt));
  MPI_Barrier(0);
  return 0;
}
static inline int MPI_Allgather(const void *sb, int sc, MPI_Datatype st, void *rb, int rc, MPI_Datatype rt, MPI_Comm c) {
  (void)c;
  long long n = (long long)sc * st; phy_shim_check(n);
  memcpy(phy_shim_arena_p->slots[phy_shim_rank], sb, (size_t)n);
  MPI_Barrier(0);
  for (int r = 0; r < phy_shim_np; ++r) memcpy((char *)rb + (long long)r * rc * rt, phy_shim_arena_p->slots[r], (size_t)((long long)rc * rt));
  MPI_Barrier(0);
  return 0;
}
/* Exscan / Allreduce over 8-byte integers or doubles only (what the host driver needs). */
static inline int MPI_Exscan(const void *sb, void *rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)c;
  if (t != 8 || op != MPI_SUM) { fprintf(stderr, "mpi shim: Exscan supports 8-byte integer SUM only\n"); _exit(97); }
  long long n = (long long)count * t; phy_shim_check(n);
  memcpy(phy_shim_arena_p->slots[phy_shim_rank], sb, (size_t)n);
  MPI_Barrier(0);
  if (phy_shim_rank > 0) {
    long long *out = (long long *)rb;
    for (int i = 0; i < count; ++i) out[i] = 0;
    for (int r = 0; r < phy_shim_rank; ++r)
      for (int i = 0; i < count; ++i) out[i] += ((long long *)phy_shim_arena_p->slots[r])[i];
  }
  MPI_Barrier(0);
  return 0;
}
static inline int MPI_Allreduce(const void *sb, void *rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)c;
  if (t != 8) { fprintf(stderr, "mpi shim: Allreduce supports MPI_DOUBLE only\n"); _exit(97); }
  long long n = (long long)count * t; phy_shim_check(n);
  memcpy(phy_shim_arena_p->slots[phy_shim_rank], sb, (size_t)n);
  MPI_Barrier(0);
  double *out = (double *)rb;
  for (int i = 0; i < count; ++i) {
    double acc = ((double *)phy_shim_arena_p->slots[0])[i];
    for (int r = 1; r < phy_shim_np; ++r) {
      double v = ((double *)phy_shim_arena_p->slots[r])[i];
      if (op == MPI_SUM) acc += v; else if (v > acc) acc = v;
    }
    out[i] = acc;
  }
  MPI_Barrier(0);
  return 0;
}

#endif /* PHY_MPI_SHIM_H */
