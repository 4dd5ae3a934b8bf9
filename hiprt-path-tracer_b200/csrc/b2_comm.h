/* b2_comm.h -- internal C++ view of the multi-GPU layer (b2_comm.cu); the C ABI is in include/b2pt.h, group 1b. */
#ifndef B2_COMM_H
#define B2_COMM_H

#include <vector>

struct b2pt_comm;

/* frame rows within `halo` rows of a band rank owns (halo = 0: the owned rows), ascending */
std::vector<int> b2_partition_rows(int height, int rank, int world, int band_rows, int halo);
/* per peer of `rank`: the rows it sends to / receives from that peer in a halo exchange (ascending; same order on both sides) */
void b2_halo_plan_rows(int height, int rank, int world, int band_rows, int halo, std::vector<int>& peers, std::vector<std::vector<int>>& send_rows,
                       std::vector<std::vector<int>>& recv_rows);
int b2_comm_rank(const b2pt_comm* c);
int b2_comm_world(const b2pt_comm* c);

#endif
