/*
 * Multi-process MPI stand-in.  TEST INFRASTRUCTURE ONLY -- see include_mp/mpi.h.
 *
 * The ranks are ordinary processes started by oracle/mprun.py with MGPSHIM_RANK / MGPSHIM_SIZE / MGPSHIM_SHM in
 * their environment; MGPSHIM_SHM names a zero-filled file under /dev/shm that every rank maps:
 *     [header][size x MGPSHIM_SLOT_MB message slots][MGPSHIM_SCRATCH_MB of scratch for the FFT stand-in]
 * Every operation the reference uses (Barrier, Allreduce, Reduce, Allgather, Gather, Bcast, Sendrecv) is called by
 * all ranks in the same order (the reference is SPMD and only uses MPI_COMM_WORLD), so each one is "publish my
 * part in my slot, barrier, read the others' slots, barrier".  Reductions add in rank order on every rank: all
 * ranks get bit-identical results.  Without the environment variables the library behaves as one rank.
 *
 * This lets the UNMODIFIED reference run as the multi-rank program it is (x-slabs, halo exchanges, particle
 * migration) on the host cores of a box that has no MPI: bench.py's CPU baseline and a multi-rank parity oracle.
 */
#include <mpi.h>

#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define MAXR 256

struct hdr {
  atomic_int count, sense;
  size_t slot_bytes, scratch_off, scratch_bytes;
  size_t msg_bytes[MAXR];
  int msg_dest[MAXR];
};

static struct hdr *H = NULL;
static char *base = NULL;
static int g_rank = 0, g_size = 1, g_sense = 0;

static void die(const char *m) {
  fprintf(stderr, "[mpi shim, rank %d] %s\n", g_rank, m);
  fflush(NULL);
  _exit(70);
}

static char *slot(int r) { return base + sizeof(struct hdr) + (size_t) r * H->slot_bytes; }

int mgp_mp_rank(void) { return g_rank; }
int mgp_mp_size(void) { return g_size; }

void mgp_mp_barrier(void) {
  if (g_size == 1) return;
  g_sense = !g_sense;
  if (atomic_fetch_add(&H->count, 1) == g_size - 1) {
    atomic_store(&H->count, 0);
    atomic_store(&H->sense, g_sense);
  } else {
    int spins = 0;
    while (atomic_load(&H->sense) != g_sense)
      if (++spins > 200) { sched_yield(); spins = 0; }
  }
}

/* scratch shared by all ranks (the FFT stand-in's full grid); private memory when there is one rank */
void *mgp_mp_scratch(size_t bytes) {
  static void *priv = NULL;
  static size_t priv_bytes = 0;
  if (g_size == 1) {
    if (bytes > priv_bytes) { free(priv); priv = malloc(bytes); priv_bytes = bytes; }
    return priv;
  }
  if (bytes > H->scratch_bytes) die("shared scratch too small (MGPSHIM_SCRATCH_MB)");
  return base + H->scratch_off;
}

int MPI_Init(int *argc, char ***argv) {
  (void) argc; (void) argv;
  const char *er = getenv("MGPSHIM_RANK"), *es = getenv("MGPSHIM_SIZE"), *ef = getenv("MGPSHIM_SHM");
  if (er && es && ef && atoi(es) > 1) {
    g_rank = atoi(er); g_size = atoi(es);
    if (g_size > MAXR || g_rank < 0 || g_rank >= g_size) die("bad MGPSHIM_RANK / MGPSHIM_SIZE");
    int fd = open(ef, O_RDWR);
    if (fd < 0) die("cannot open MGPSHIM_SHM");
    struct stat st;
    if (fstat(fd, &st)) die("fstat");
    base = (char *) mmap(NULL, (size_t) st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (base == MAP_FAILED) die("mmap");
    close(fd);
    H = (struct hdr *) base;            /* the launcher created the file zero-filled: barrier state starts at 0 */
    const char *sm = getenv("MGPSHIM_SLOT_MB"), *cm = getenv("MGPSHIM_SCRATCH_MB");
    const size_t slot_b = (size_t) (sm ? atol(sm) : 64) << 20, scr_b = (size_t) (cm ? atol(cm) : 64) << 20;
    /* every rank writes the same three values: no ordering needed */
    H->slot_bytes = slot_b;
    H->scratch_off = (sizeof(struct hdr) + (size_t) g_size * slot_b + 4095) / 4096 * 4096;
    H->scratch_bytes = scr_b;
    if (H->scratch_off + scr_b > (size_t) st.st_size) die("segment too small");
  } else {
    g_rank = 0; g_size = 1;
    H = (struct hdr *) calloc(1, sizeof(struct hdr) + (1 << 20));
    base = (char *) H;
    H->slot_bytes = 1 << 20;
  }
  return MPI_SUCCESS;
}
int MPI_Finalize(void) { mgp_mp_barrier(); return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) {
  (void) comm;
  fprintf(stderr, "[mpi shim, rank %d] MPI_Abort(%d)\n", g_rank, code);
  fflush(NULL);
  _exit(code ? code : 1);               /* the launcher kills the other ranks */
}
int MPI_Barrier(MPI_Comm comm) { (void) comm; mgp_mp_barrier(); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void) comm; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void) comm; *size = g_size; return MPI_SUCCESS; }

/* out[i] = op(out[i], in[i]) */
static void combine(void *out, const void *in, size_t n, MPI_Datatype t, MPI_Op op) {
  const int kind = MGP_DT_KIND(t);
  const size_t sz = MGP_DT_SIZE(t);
#define LOOP(T)                                                                  \
  do {                                                                           \
    T *o = (T *) out; const T *a = (const T *) in;                               \
    for (size_t i = 0; i < n; i++) {                                             \
      if (op == MPI_SUM) o[i] = (T) (o[i] + a[i]);                               \
      else if (op == MPI_MAX) { if (a[i] > o[i]) o[i] = a[i]; }                  \
      else { if (a[i] < o[i]) o[i] = a[i]; }                                     \
    }                                                                            \
  } while (0)
  if (kind == 3 && sz == 8) LOOP(double);
  else if (kind == 3 && sz == 4) LOOP(float);
  else if (kind == 1 && sz == 4) LOOP(int);
  else if (kind == 1 && sz == 8) LOOP(long long);
  else if (kind == 2 && sz == 4) LOOP(unsigned);
  else if (kind == 2 && sz == 8) LOOP(unsigned long long);
  else die("reduction on an unsupported datatype");
#undef LOOP
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm) {
  (void) comm;
  const size_t sz = MGP_DT_SIZE(t);
  if (g_size == 1) {
    if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memmove(recvbuf, sendbuf, (size_t) count * sz);
    return MPI_SUCCESS;
  }
  const char *src = (const char *) (sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf);
  const size_t per = H->slot_bytes / sz;
  for (size_t done = 0; done < (size_t) count || done == 0; done += per) {
    const size_t n = (size_t) count - done < per ? (size_t) count - done : per;
    memcpy(slot(g_rank), src + done * sz, n * sz);
    mgp_mp_barrier();
    char *dst = (char *) recvbuf + done * sz;
    memcpy(dst, slot(0), n * sz);
    for (int r = 1; r < g_size; r++) combine(dst, slot(r), n, t, op);
    mgp_mp_barrier();
    if (count == 0) break;
  }
  return MPI_SUCCESS;
}
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm) {
  /* every rank computes the result; only the root's recvbuf is defined by MPI, the others get a private copy */
  if (g_size > 1 && g_rank != root) {
    void *tmp = malloc((size_t) count * MGP_DT_SIZE(t) + 16);
    const void *src = sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf;
    memcpy(tmp, src, (size_t) count * MGP_DT_SIZE(t));
    MPI_Allreduce(MPI_IN_PLACE, tmp, count, t, op, comm);
    free(tmp);
    return MPI_SUCCESS;
  }
  return MPI_Allreduce(sendbuf, recvbuf, count, t, op, comm);
}
int MPI_Allgather(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, int rcount, MPI_Datatype rt, MPI_Comm comm) {
  (void) comm;
  const size_t rb = (size_t) rcount * MGP_DT_SIZE(rt);
  const size_t sb = sendbuf == MPI_IN_PLACE ? rb : (size_t) scount * MGP_DT_SIZE(st);
  const char *src = sendbuf == MPI_IN_PLACE ? (const char *) recvbuf + (size_t) g_rank * rb : (const char *) sendbuf;
  if (g_size == 1) {
    if (src != recvbuf) memmove(recvbuf, src, sb);
    return MPI_SUCCESS;
  }
  if (sb > H->slot_bytes) die("MPI_Allgather block larger than a message slot");
  memcpy(slot(g_rank), src, sb);
  mgp_mp_barrier();
  for (int r = 0; r < g_size; r++) memcpy((char *) recvbuf + (size_t) r * rb, slot(r), rb < sb ? rb : sb);
  mgp_mp_barrier();
  return MPI_SUCCESS;
}
int MPI_Gather(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, int rcount, MPI_Datatype rt, int root, MPI_Comm comm) {
  if (g_size > 1 && g_rank != root) {    /* recvbuf is only significant at the root */
    const size_t sb = (size_t) scount * MGP_DT_SIZE(st);
    void *tmp = malloc(sb * (size_t) g_size + 16);
    MPI_Allgather(sendbuf, scount, st, tmp, scount, st, comm);
    free(tmp);
    return MPI_SUCCESS;
  }
  return MPI_Allgather(sendbuf, scount, st, recvbuf, rcount, rt, comm);
}
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm) {
  (void) comm;
  if (g_size == 1) return MPI_SUCCESS;
  const size_t sz = MGP_DT_SIZE(t), per = H->slot_bytes / sz;
  for (size_t done = 0; done < (size_t) count; done += per) {
    const size_t n = (size_t) count - done < per ? (size_t) count - done : per;
    if (g_rank == root) memcpy(slot(root), (char *) buf + done * sz, n * sz);
    mgp_mp_barrier();
    if (g_rank != root) memcpy((char *) buf + done * sz, slot(root), n * sz);
    mgp_mp_barrier();
  }
  return MPI_SUCCESS;
}
int MPI_Sendrecv(const void *sendbuf, int scount, MPI_Datatype st, int dest, int stag,
                 void *recvbuf, int rcount, MPI_Datatype rt, int source, int rtag,
                 MPI_Comm comm, MPI_Status *status) {
  (void) stag; (void) comm;
  const size_t sb = dest == MPI_PROC_NULL ? 0 : (size_t) scount * MGP_DT_SIZE(st);
  const size_t rb = source == MPI_PROC_NULL ? 0 : (size_t) rcount * MGP_DT_SIZE(rt);
  if (g_size == 1) {
    if (dest == MPI_PROC_NULL || source == MPI_PROC_NULL) return MPI_SUCCESS;
    const size_t n = sb < rb ? sb : rb;
    if (n && sendbuf != recvbuf) memmove(recvbuf, sendbuf, n);
    if (status) { status->MPI_SOURCE = 0; status->MPI_TAG = rtag; status->MPI_ERROR = MPI_SUCCESS; }
    return MPI_SUCCESS;
  }
  /* all ranks are inside a Sendrecv of the same exchange: publish sizes, then move the payloads slot by slot */
  H->msg_bytes[g_rank] = sb;
  H->msg_dest[g_rank] = dest;
  mgp_mp_barrier();
  size_t maxb = 0;
  for (int r = 0; r < g_size; r++) if (H->msg_bytes[r] > maxb) maxb = H->msg_bytes[r];
  size_t from_src = 0;
  if (source != MPI_PROC_NULL) {
    if (H->msg_dest[source] != g_rank) die("MPI_Sendrecv: the source rank is not sending to this rank");
    from_src = H->msg_bytes[source];
    if (from_src > rb) die("MPI_Sendrecv: message longer than the receive buffer");
  }
  const size_t S = H->slot_bytes;
  for (size_t off = 0; off < maxb || off == 0; off += S) {
    if (off < sb) memcpy(slot(g_rank), (const char *) sendbuf + off, sb - off < S ? sb - off : S);
    mgp_mp_barrier();
    if (off < from_src) memcpy((char *) recvbuf + off, slot(source), from_src - off < S ? from_src - off : S);
    mgp_mp_barrier();
    if (maxb == 0) break;
  }
  if (status) { status->MPI_SOURCE = source; status->MPI_TAG = rtag; status->MPI_ERROR = MPI_SUCCESS; }
  return MPI_SUCCESS;
}
int MPI_Type_match_size(int typeclass, int size, MPI_Datatype *t) {
  (void) typeclass; *t = MGP_DT(3, size); return MPI_SUCCESS;
}
int MPI_Type_create_struct(int count, const int *blocklengths, const MPI_Aint *offsets,
                           const MPI_Datatype *types, MPI_Datatype *newtype) {
  /* extent = furthest member end, rounded up to the widest member alignment */
  long end = 0, align = 1;
  for (int i = 0; i < count; i++) {
    const long sz = (long) MGP_DT_SIZE(types[i]);
    long e = (long) offsets[i] + (long) blocklengths[i] * sz;
    if (e > end) end = e;
    if (sz > align) align = sz;
  }
  end = (end + align - 1) / align * align;
  *newtype = MGP_DT(0, end);
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *t) { (void) t; return MPI_SUCCESS; }

int MPI_Type_struct(int count, int *blocklengths, MPI_Aint *offsets, MPI_Datatype *types, MPI_Datatype *newtype) {
  return MPI_Type_create_struct(count, blocklengths, offsets, types, newtype);
}
/* every rank learns every count, then rank r's piece travels like a broadcast from r; only the root keeps it */
int MPI_Gatherv(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, const int *rcounts, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm comm) {
  (void) rcounts; (void) rt;
  const size_t sz = MGP_DT_SIZE(st);
  int *counts = (int *) malloc(sizeof(int) * (size_t) g_size);
  MPI_Allgather(&scount, 1, MPI_INT, counts, 1, MPI_INT, comm);
  for (int r = 0; r < g_size; r++) {
    if (counts[r] == 0) continue;
    char *tmp = (g_rank == root) ? (char *) recvbuf + (size_t) displs[r] * sz : (char *) malloc((size_t) counts[r] * sz);
    if (g_rank == r) memcpy(tmp, sendbuf, (size_t) counts[r] * sz);
    MPI_Bcast(tmp, (int) ((size_t) counts[r] * sz), MPI_BYTE, r, comm);
    if (g_rank != root) free(tmp);
  }
  free(counts);
  return MPI_SUCCESS;
}
