/* oracle/shims/mpi.h -- TEST INFRASTRUCTURE ONLY.
 * Single-rank stand-in for the MPI calls the reference makes (exec/boltz.c:34-36,98,180-181,
 * 363-385,422; src/transportroutines.c:100-164,249-343; src/output.c:242-243,352,400;
 * src/mesh_setup.c:43-44; src/restart.c:16,61). No MPI exists in this image. With one rank no
 * Send/Recv is ever reached; they abort if called. */
#ifndef ORC_SHIM_MPI_H
#define ORC_SHIM_MPI_H
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_SUCCESS 0
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm c, int *size);
int MPI_Comm_rank(MPI_Comm c, int *rank);
int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s);
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Barrier(MPI_Comm c);
double MPI_Wtime(void);
#endif
