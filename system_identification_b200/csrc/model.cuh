// Device-side model image and compile-time envelope shared by all kernels.
#pragma once
#include <cstdint>

namespace sysid {

constexpr int MAXJ = 14;     // joints including the universe (index 0)
constexpr int MAXB = 13;     // bodies = moving joints (free-flyer root + 12 revolute)
constexpr int MAXV = 18;     // nv = 6 + 12
constexpr int MAXD = 12;     // actuated joints
constexpr int MAXEE = 4;     // contact frames
constexpr int MAXCH = 6;     // longest foot chain (joints between the foot and the root, root excluded)
constexpr int CW = 160;      // padded width of one stacked row: 10*nb + 2*d + 1 (tau column) <= 155 -> 160
#ifndef SYSID_ST_GROUP
#define SYSID_ST_GROUP 2
#endif
constexpr int ST_GROUP = SYSID_ST_GROUP;   // bodies of a leg one tile-fill task emits (the walk down the leg is shared between them)
constexpr int ST_MAXLEG = 4;  // legs (simple chains below the root) the structured-basis Gram kernel handles
constexpr int PROJ_PARTS = 4;      // the projection phase of the fused kernel deals every sample's bodies to this many warps
constexpr int PROJ_MAXITEMS = 12;  // chain-walk steps of one part (prefix joints it only accumulates + joints whose body it emits)

enum { JT_FF = 0, JT_RX = 1, JT_RY = 2, JT_RZ = 3, JT_RU = 4 };

// Passed to kernels BY VALUE as a __grid_constant__ parameter (lives in the constant bank, ~2 KB):
// no global symbol, so any number of models can be in flight on any streams.
struct DevModel {
    int32_t njoints, nb, nv, nq, nd, n_ee;
    int32_t nparams;                 // 10 * nb
    int32_t parent[MAXJ];
    int32_t jtype[MAXJ];
    double axis[MAXJ][3];
    double pR[MAXJ][9];              // joint placement rotation, row-major
    double pp[MAXJ][3];              // joint placement translation
    double gravity[3];
    // contact frames: chain[k][e] = e-th joint walking from the foot's joint towards the root (root excluded)
    int32_t ee_joint[MAXEE];
    double ee_off[MAXEE][3];
    int32_t chain_len[MAXEE];
    int32_t chain[MAXEE][MAXCH];
    int32_t nshared[MAXEE][MAXEE];   // length of the common suffix of two foot chains
    uint8_t colperm[CW];             // per-sample debug kernels: visiting order of the padded columns, sorted by chain depth
    // leaf chains for the chain phase of the fused kernels: fch[c][e] = e-th joint walking from the ROOT towards leaf c
    // (root excluded); joints [fch_own[c], fch_len[c]) are written by chain c, the shared prefix by an earlier chain
    int32_t nfch;
    int32_t fch_len[MAXD];
    int32_t fch_own[MAXD];
    int32_t fch[MAXD][MAXCH];
    int32_t fill_split;              // bodies of a chain are dealt round-robin to this many lane groups of the tile fill
    uint32_t submask[MAXJ];          // bit i set: body (joint) i lies in the subtree of joint j, j itself included (evaluation pass)
    // projection phase (phase_proj_mma): part p walks the joints proj_item[p][0 .. proj_n[p]); flag bit 0 = emit the body's ten columns,
    // bit 1 = first joint of a walk (restart from the base rows); proj_root[p]: the part emits the root body; proj_tail[p]:
    // mask of the friction / torque / padding 8-column groups (from column nparams on) the part emits
    int8_t proj_n[PROJ_PARTS];
    uint32_t proj_item[PROJ_PARTS][PROJ_MAXITEMS];  // joint | flags << 8 (flag bit 0: emit, bit 1: restart)
    uint8_t proj_root[PROJ_PARTS];
    uint32_t proj_tail[PROJ_PARTS];
    uint8_t proj_tailks[2][32];      // [friction][column group]: which 4-joint k-steps of the group are structurally non-zero
    // structured null-space basis (gram_struct.cuh): every child of the root heads a simple chain ("leg") with at most one contact
    // frame; the legs are dealt to two column classes of <= 6 joints.  Tile columns: class X at 72 X + [10 slot .. | 60 + slot
    // (viscous) | 66 + slot (Coulomb)], torque column 144, root body 146..155.
    int8_t st_ok;                    // 1: the model satisfies the above and the structured kernel serves it
    int8_t st_nslot[2];              // joints of class A / B
    int8_t st_slotjoint[2][6];       // class slot -> joint, -1: unused
    int8_t st_cfoot[ST_MAXLEG];      // contact frame on leg c (index into ee_*), -1: none
    int8_t st_ccls[ST_MAXLEG];       // column class of leg c
    uint32_t st_jrec[MAXJ];          // per joint >= 2: leg | position in the leg << 2 | class << 5 | class slot << 6 | body column << 9 | viscous column << 17
    // tile fill: one warp task = a run of consecutive bodies of one leg (ST_GROUP at most) for all rows of a tile:
    // kind | first joint << 4 | bodies << 8 | class slot of the first body << 12 | (class task that also writes the torque block) << 16;
    // kind 0: dense rows, 1: dense rows x root body + torque column, 2 / 3: class A / B rows; sorted by cost, heaviest first
    int8_t st_ntask;
    int8_t st_maxlen;                // joints of the longest leg
    int8_t st_nred;                  // columns of the reduced contact Jacobian: 6 + 3 (legs with a contact frame)
    uint32_t st_task[32];
};

}  // namespace sysid
