/*
 * gomc_b200.h -- C ABI of the B200-native GOMC energy/force engine.
 *
 * GOMC has no plugin API: its GPU seam is the set of free C++ functions
 * Call*GPU(VariablesCUDA*, ...) declared in src/GPU/*.cuh and called from the
 * `#ifdef GOMC_CUDA` blocks of CalculateEnergy.cpp / Ewald.cpp (SURVEY.md
 * section 8b).  This header is the flattened, C-linkage replacement of that seam:
 * plain pointers and sizes only.  Every entry point names the reference
 * declaration it replaces.  INTEGRATION.md shows the glue a GOMC maintainer
 * adds under `#ifdef GOMC_CUDA`.
 *
 * Differences from the reference seam, all deliberate:
 *  - STATEFUL: topology, charges, kinds, force-field tables, k-vectors and
 *    the structure-factor sums live on the device; coordinates are uploaded
 *    with gomcb200_set_coords / gomcb200_set_molecule_coords and stay
 *    resident.  The gomcb200_call_* entry points keep the reference's
 *    "host buffers in, scalars out" convention for a literal drop-in.
 *  - The engine bins atoms into cells on the device; the host CSR cell list
 *    of the reference (cellVector / cellStartIndex / neighborList arguments
 *    of CallBoxInterGPU) is not needed.  Box membership is given once per
 *    change with gomcb200_set_box_molecules.
 *  - All reductions are fixed-order: two calls on the same state return
 *    bit-identical results (the reference uses atomicAdd).
 *  - Errors: every function returns 0 on success or a negative GOMCB200_E*
 *    code; gomcb200_last_error() returns the message.  (The reference prints
 *    and exit()s, src/GPU/VariablesCUDA.cuh:20-39; the C++ host mirror in
 *    gomc_b200/host keeps that behaviour.)  There is no CPU fallback: with no
 *    usable CUDA device gomcb200_create fails.
 *
 * Arithmetic is IEEE double throughout.  A fractional molecule (lambda < 1, soft-core
 * pair functors and lambda-scaled Ewald terms) is set with gomcb200_update_lambda.
 * Potentials: VDW (Mie), SHIFT, SWITCH, SWITCH+Martini, EXP6.
 * Orthogonal (gomcb200_set_box_axes) and non-orthogonal (gomcb200_set_box_cell_basis) cells.
 * Threading: one host thread per engine, like the reference.
 */
#ifndef GOMC_B200_H
#define GOMC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gomcb200_engine gomcb200_engine;

enum {
  GOMCB200_OK = 0,
  GOMCB200_EINVAL = -1,  /* bad argument / call order */
  GOMCB200_ECUDA = -2,   /* CUDA runtime error (message has the details) */
  GOMCB200_ENODEV = -3,  /* no usable sm_100 device */
  GOMCB200_ENOMEM = -4,
  GOMCB200_EKMAX = -5    /* more k-vectors than imageTotal
                            (src/Ewald.cpp:898-902 "Kmax exceeded") */
};

/* src/GPU/ConstantDefinitionsCUDAKernel.cuh:18-20 */
enum {
  GOMCB200_VDW_STD = 0,
  GOMCB200_VDW_SHIFT = 1,
  GOMCB200_VDW_SWITCH = 2, /* with isMartini: FF_SWITCH_MARTINI */
  GOMCB200_VDW_EXP6 = 3
};

/* which force / k-vector buffer */
enum {
  GOMCB200_ATOM_FORCE = 0,     /* System::atomForceRef    */
  GOMCB200_MOL_FORCE = 1,      /* System::molForceRef     */
  GOMCB200_ATOM_FORCE_REC = 2, /* System::atomForceRecRef */
  GOMCB200_MOL_FORCE_REC = 3,  /* System::molForceRecRef  */
  GOMCB200_MOL_TORQUE = 4      /* MultiParticle molTorque */
};
enum { GOMCB200_K_NEW = 0, GOMCB200_K_REF = 1,           /* kx[]  vs kxRef[]   */
       GOMCB200_K_DEVICE = 2 };  /* OR-ed in: read back the device-resident copy */
enum { GOMCB200_SUM_NEW = 0, GOMCB200_SUM_REF = 1 };   /* sumRnew vs sumRref */

const char *gomcb200_last_error(void);
int gomcb200_version(void);

/* ---- lifecycle --------------------------------------------------------- */
/* Replaces `new VariablesCUDA()` (src/FFParticle.cpp:56) and the device pick
 * in src/Main.cpp:286-327.  device < 0 -> current device.  nBoxes = BOX_TOTAL
 * (1 or 2). */
int gomcb200_create(gomcb200_engine **out, int device, int nBoxes);
/* DestroyCUDAVars / DestroyEwaldCUDAVars,
 * src/GPU/ConstantDefinitionsCUDAKernel.cuh:43-46 */
int gomcb200_destroy(gomcb200_engine *e);
/* Number of kernels this engine has launched so far (bench.py gpu_launches). */
long long gomcb200_launch_count(const gomcb200_engine *e);

/* InitGPUForceField, src/GPU/ConstantDefinitionsCUDAKernel.cuh:24-30.
 * Tables are [count*count], index kind1 + kind2*count (src/FFParticle.h:111).
 * rCutCoulomb / alpha are per box [nBoxes]. */
int gomcb200_init_forcefield(gomcb200_engine *e, const double *sigmaSq,
                             const double *epsilon_cn, const double *n,
                             int vdwKind, int isMartini, int count,
                             double rCut, const double *rCutCoulomb,
                             double rCutLow, double rOn, const double *alpha,
                             int ewald, int electrostatic,
                             double diElectric_1);

/* Soft-core constants of the free-energy / NeMTMC functors (forcefield.sc_alpha,
 * sc_sigma_6, sc_power, sc_coul; src/Forcefield.cpp:58-75) -- what the reference
 * passes with every CallBoxInterGPU / CallBoxForceGPU call. */
int gomcb200_init_softcore(gomcb200_engine *e, double sc_alpha, double sc_sigma_6,
                           int sc_power, int sc_coul);
/* UpdateGPULambda (ConstantDefinitionsCUDAKernel.cuh:31-33; lib/Lambda.h:65-88):
 * the fractional molecule of a box.  Pair terms use lambdaVDW / lambdaCoulomb for
 * pairs involving it (CalculateEnergy::GetLambdaVDW, src/CalculateEnergy.cpp:
 * 1558-1572); every Ewald term sees its charges times sqrt(lambdaCoulomb)
 * (Ewald::GetLambdaCoef, src/Ewald.cpp:1598-1602).  molKind is the molecule's kind
 * index (BoxSelf compares it with the molecule index, :1140-1155).
 * isFraction = 0 clears the state. */
int gomcb200_update_lambda(gomcb200_engine *e, int box, int molIndex, int molKind,
                           double lambdaVDW, double lambdaCoulomb, int isFraction);
/* InitExp6VariablesCUDA, src/GPU/ConstantDefinitionsCUDAKernel.cuh:34-35:
 * the rMin / expConst / rMaxSq tables FF_EXP6::Init derives with Brent's method
 * (src/FFExp6.h:99-147); required before any energy call when vdwKind is EXP6. */
int gomcb200_init_exp6(gomcb200_engine *e, const double *rMin,
                       const double *expConst, const double *rMaxSq, int size);

/* InitCoordinatesCUDA (ConstantDefinitionsCUDAKernel.cuh:31-32) plus the
 * per-atom vectors CalculateEnergy::Init / Ewald::Init build
 * (src/CalculateEnergy.cpp:60-81, src/Ewald.cpp:100-127).
 * molStart has nMols+1 entries (src/Molecules.h:46-51). */
int gomcb200_init_topology(gomcb200_engine *e, int nAtoms, int nMols,
                           const int *particleKind, const int *particleMol,
                           const double *particleCharge, const int *molStart);

/* MoleculeLookup box list (molLookup.BoxBegin(box)..BoxEnd(box)); call again
 * after an accepted molecule transfer. */
int gomcb200_set_box_molecules(gomcb200_engine *e, int box,
                               const int *molIndices, int nMolsInBox);

/* UpdateCellBasisCUDA (ConstantDefinitionsCUDAKernel.cuh:38-39) for an
 * orthogonal cell: axis = BoxDimensions::axis.Get(box). */
int gomcb200_set_box_axes(gomcb200_engine *e, int box, const double axis[3]);

/* UpdateCellBasisCUDA + UpdateInvCellBasisCUDA (ConstantDefinitionsCUDAKernel.cuh:38-42)
 * for a non-orthogonal cell: row-major 3x3, rows = the NORMALISED cell vectors
 * (BoxDimensionsNonOrth::cellBasis[box]) and the inverse of that matrix
 * (cellBasis_Inv[box]); axis = the cell edge lengths.  NULL matrices restore the
 * orthogonal case.  Slanted boxes use the per-term reciprocal kernels (the valid
 * c range of an (a,b) row is not symmetric there); the pair path is unchanged. */
int gomcb200_set_box_cell_basis(gomcb200_engine *e, int box,
                                const double cellBasis[9],
                                const double cellBasisInv[9],
                                const double axis[3]);

/* Coordinates / centres of mass (System::coordinates, System::com), global
 * atom / molecule indexing; [first, first+count). */
int gomcb200_set_coords(gomcb200_engine *e, const double *x, const double *y,
                        const double *z, int first, int count);
int gomcb200_get_coords(gomcb200_engine *e, double *x, double *y, double *z,
                        int first, int count);
int gomcb200_set_com(gomcb200_engine *e, const double *x, const double *y,
                     const double *z, int first, int count);
/* Molecule centres of mass of the active coordinate set (device -> host). */
int gomcb200_get_com(gomcb200_engine *e, double *x, double *y, double *z,
                     int first, int count);
/* Accepted single-molecule move: new coordinates + COM of one molecule
 * (what Translate::Accept copies, src/moves/Translate.h:106-113). */
int gomcb200_set_molecule_coords(gomcb200_engine *e, int molIndex,
                                 const double *x, const double *y,
                                 const double *z, const double com[3]);

/* ---- pair path (CalculateEnergy) --------------------------------------- */
/* CallBoxInterGPU, src/GPU/CalculateEnergyCUDAKernel.cuh:17-26 ==
 * CalculateEnergy::BoxInter pair sums, src/CalculateEnergy.cpp:157-266. */
int gomcb200_box_inter(gomcb200_engine *e, int box, double *LJEn,
                       double *REn);
/* CallBoxForceGPU, src/GPU/CalculateForceCUDAKernel.cuh:17-30 ==
 * CalculateEnergy::BoxForce, src/CalculateEnergy.cpp:268-406.  Writes the
 * resident ATOM_FORCE / MOL_FORCE buffers (reset for this box first, as
 * ResetForce does); read them with gomcb200_get_forces. */
int gomcb200_box_force(gomcb200_engine *e, int box, double *LJEn,
                       double *REn);
/* CalculateEnergy::MoleculeInter, src/CalculateEnergy.cpp:581-686 (host-only
 * in the reference GPU build).  The molecule is excluded from its own
 * neighbourhood, which is what CellList::RemoveMol achieves. */
int gomcb200_molecule_inter(gomcb200_engine *e, int box, int molIndex,
                            const double *newX, const double *newY,
                            const double *newZ, double *dLJ, double *dReal,
                            int *overlap);
/* One single-molecule trial = MoleculeInter + MolReciprocal (what Translate::CalcEn /
 * Rotate::CalcEn call back to back, src/moves/Translate.h:82-95) queued together
 * with a single host synchronisation.  The reciprocal term is computed even when
 * the move overlaps (the reference skips it then; the caller ignores it). */
int gomcb200_molecule_trial(gomcb200_engine *e, int box, int molIndex,
                            const double *newX, const double *newY,
                            const double *newZ, double *dLJ, double *dReal,
                            int *overlap, double *energyRecipNew);
/* CalculateEnergy::ParticleInter, src/CalculateEnergy.cpp:727-785: en[] and
 * real[] are incremented, overlap[] or-ed. */
int gomcb200_particle_inter(gomcb200_engine *e, int box, int molIndex,
                            int partIndex, int trials, const double *tx,
                            const double *ty, const double *tz, double *en,
                            double *real, int *overlap);
/* CalculateEnergy::ParticleNonbonded, src/CalculateEnergy.cpp:689-725 (called next to
 * ParticleInter by every CBMC growth step, e.g. src/cbmc/DCSingle.cpp:48,86): the 1-N
 * intramolecular non-bonded energy (FFParticle::CalcEn + CalcCoulombAdd_1_4 with NB = true)
 * of `trials` trial positions of one site of kind kindI / charge chargeI against the partner
 * sites that already exist in the trial molecule -- the caller walks kind.sortedNB(partIndex)
 * and keeps the entries with TrialMol::AtomExists, in that order.  inter[] is incremented. */
int gomcb200_particle_nonbonded(gomcb200_engine *e, int box, int kindI, double chargeI,
                                int nPartners, const int *partnerKind,
                                const double *partnerCharge, const double *px,
                                const double *py, const double *pz, int trials,
                                const double *tx, const double *ty, const double *tz,
                                double *inter);
/* CalculateEnergy::CalculateTorque, src/CalculateEnergy.cpp:1365-1406, from
 * the resident ATOM_FORCE + ATOM_FORCE_REC buffers and COM. */
int gomcb200_calculate_torque(gomcb200_engine *e, int box);
int gomcb200_get_forces(gomcb200_engine *e, int which, double *x, double *y,
                        double *z, int first, int count);

/* ---- Ewald reciprocal path --------------------------------------------- */
/* InitEwaldVariablesCUDA, ConstantDefinitionsCUDAKernel.cuh:33 (capacity of
 * the per-box k arrays, Ewald::AllocMem src/Ewald.cpp:141-188).
 * recip_rcut[nBoxes] = Forcefield::recip_rcut (src/Forcefield.cpp:82); it is
 * passed explicitly so that the k-vector membership test is bit-identical to
 * the host's. */
int gomcb200_init_ewald(gomcb200_engine *e, int imageTotal,
                        const double *recip_rcut);
/* Ewald::RecipInit (RecipInitOrth, src/Ewald.cpp:847-903) for the given
 * axes into the NEW k set (kx[box]...); returns imageSize[box] and kmax. */
int gomcb200_recip_init(gomcb200_engine *e, int box, const double axis[3],
                        int *imageSize, int *kmax);
/* The same for the dimensions of a volume trial (VolumeTransfer::CalcEn,
 * src/moves/VolumeTransfer.h:160-189: RecipInit(box, newDim)).  volume =
 * newDim.volume[box]: BoxDimensions::SetVolume (src/BoxDimensions.cpp:214-226) stores
 * oldVolume + delta and cbrt-scales the axes, and RecipInitOrth's prefactor divides by
 * that stored value (src/Ewald.cpp:857), which differs from the product of the axes in
 * the last bit.  volume == 0: the product (gomcb200_recip_init). */
int gomcb200_recip_init_volume(gomcb200_engine *e, int box, const double axis[3],
                               double volume, int *imageSize, int *kmax);
/* Ewald::RecipCountInit, src/Ewald.cpp:968-1018 (count only; excess = the
 * ensemble head-room factor 1.0 / 1.25 / 1.5). */
int gomcb200_recip_count(gomcb200_engine *e, int box, const double axis[3],
                         double excess, int *imageSize);
int gomcb200_get_kvectors(gomcb200_engine *e, int box, int which, double *kx,
                          double *ky, double *kz, double *hsqr,
                          double *prefact, int n);
/* CallBoxReciprocalSetupGPU (CalculateEwaldCUDAKernel.cuh:27-33) ==
 * Ewald::BoxReciprocalSetup, src/Ewald.cpp:193-274: NEW k set. */
int gomcb200_box_reciprocal_setup(gomcb200_engine *e, int box,
                                  double *energyRecip);
/* CallBoxReciprocalSumsGPU (CalculateEwaldCUDAKernel.cuh:35-38) ==
 * Ewald::BoxReciprocalSums, src/Ewald.cpp:281-361: REF k set. */
int gomcb200_box_reciprocal_sums(gomcb200_engine *e, int box,
                                 double *energyRecip);
/* Ewald::BoxReciprocal, src/Ewald.cpp:375-406. */
int gomcb200_box_reciprocal(gomcb200_engine *e, int box, int isNewVolume,
                            double *energyRecip);
/* CallMolReciprocalGPU (CalculateEwaldCUDAKernel.cuh:40-44) ==
 * Ewald::MolReciprocal, src/Ewald.cpp:409-473; old coordinates are the
 * resident ones.  Returns E_new (caller subtracts sysPotRef recip). */
int gomcb200_mol_reciprocal(gomcb200_engine *e, int box, int molIndex,
                            const double *newX, const double *newY,
                            const double *newZ, double *energyRecipNew);
/* CallSwapReciprocalGPU (CalculateEwaldCUDAKernel.cuh:54-57) ==
 * Ewald::SwapDestRecip (insert=1, src/Ewald.cpp:478-531) /
 * SwapSourceRecip (insert=0, :657-710).  Charges are those of molIndex. */
int gomcb200_swap_reciprocal(gomcb200_engine *e, int box, int molIndex,
                             const double *x, const double *y,
                             const double *z, int insert,
                             double *energyRecipNew);
/* CallBoxForceReciprocalGPU (CalculateEwaldCUDAKernel.cuh:16-25) ==
 * Ewald::BoxForceReciprocal, src/Ewald.cpp:1496-1596 incl. the
 * intramolecular correction force; writes ATOM_FORCE_REC / MOL_FORCE_REC. */
int gomcb200_box_force_reciprocal(gomcb200_engine *e, int box);
/* Lazy D2H of the structure-factor sums for host-only reference callers
 * (MolExchangeReciprocal src/Ewald.cpp:794-806, ChangeRecip :626-630). */
int gomcb200_get_recip_sums(gomcb200_engine *e, int box, int which,
                            double *sumR, double *sumI, int n);
/* ---- MultiParticle move (device-resident trial state) -------------------- */
/* CallTranslateParticlesGPU (moveType 0) / CallRotateParticlesGPU (moveType 1),
 * src/GPU/TransformParticlesCUDAKernel.cuh:20-37 == MultiParticle::
 * CalculateTrialDistRot, src/moves/MultiParticle.h:566-715.  Reads the resident
 * reference coordinates, COMs and molecule forces (+ reciprocal) or torques,
 * draws Random123Wrapper's variates (Philox4x64-10, counter {molecule, key},
 * key {step, seed}) and writes the trial coordinates / COMs, t_k or r_k and the
 * inForceRange flags into the engine's second coordinate set.
 * isMoleculeInvolved: nMols flags, or NULL for every molecule of the box. */
int gomcb200_mp_transform(gomcb200_engine *e, int box, int moveType, double max,
                          double lambdaBETA, unsigned long long step,
                          unsigned int key, unsigned long long seed,
                          const signed char *isMoleculeInvolved);
/* BrownianMotionTranslateParticlesGPU / BrownianMotionRotateParticlesGPU
 * (src/GPU/TransformParticlesCUDAKernel.cuh:40-52) == MultiParticleBrownian::
 * CalculateTrialDistRot, src/moves/MultiParticleBrownianMotion.h: k = force*BETA*max
 * + N(0, sqrt(2 max)) (Box-Muller on the same Philox stream), every molecule moved.
 * gomcb200_bm_coeff is MultiParticleBrownian::GetCoeff (the LOG of the weight
 * ratio).  Trial state handling as for gomcb200_mp_transform. */
int gomcb200_bm_transform(gomcb200_engine *e, int box, int moveType, double max,
                          double BETA, unsigned long long step, unsigned int key,
                          unsigned long long seed,
                          const signed char *isMoleculeInvolved);
int gomcb200_bm_coeff(gomcb200_engine *e, int box, int moveType, double max,
                      double BETA, double *wRatio);
/* t_k / r_k (nMols each) and inForceRange (nMols) of the last transform. */
int gomcb200_mp_get_trial(gomcb200_engine *e, double *kx, double *ky, double *kz,
                          int *inForceRange);
/* trial = 1: the trial coordinates/COMs become the active set and the force
 * buffers switch to the "New" set (MultiParticle::CalcEn works on newMolsPos,
 * atomForceNew, ..., :414-441); trial = 0: back to the reference (reject).
 * O(1) pointer exchange. */
int gomcb200_mp_select(gomcb200_engine *e, int trial);
/* MultiParticle::GetCoeff, :460-513, from the reference and new force sets. */
int gomcb200_mp_coeff(gomcb200_engine *e, int box, int moveType, double max,
                      double lambdaBETA, double *wRatio);
/* MultiParticle::Accept, :522-534: the active trial set becomes the reference
 * (coordinates, COMs, forces, torques) and UpdateRecip(box). */
int gomcb200_mp_accept(gomcb200_engine *e, int box);

/* CallBoxInterForceGPU (CalculateForceCUDAKernel.cuh:17-34) == the pair part of
 * CalculateEnergy::VirialCalc, src/CalculateEnergy.cpp:411-567: diagonal of the LJ
 * virial tensor vT and of the real-space Coulomb tensor rT (qqFact included); the
 * reference computes only the diagonal.  Uses the resident coordinates and COMs.
 * The tail correction (VirialCorrection, :1317-1336) stays a host formula. */
int gomcb200_box_inter_virial(gomcb200_engine *e, int box, double vT[3],
                              double rT[3]);
/* CallVirialReciprocalGPU (CalculateForceCUDAKernel.cuh:45-50) ==
 * Ewald::VirialReciprocal, src/Ewald.cpp:1168-1305: diagonal wT from the
 * reference sums and Ref k-vectors. */
int gomcb200_virial_reciprocal(gomcb200_engine *e, int box, double wT[3]);
/* Ewald::MolExchangeReciprocal (src/Ewald.cpp:714-826; the reference's
 * CallMolExchangeReciprocalGPU, CalculateEwaldCUDAKernel.cuh:45-48, only uploads
 * host results -- here the sums are computed on the device).  n weighted point
 * charges in the reference's loop order: w = +q*lambdaCoef for every charged atom
 * of the inserted molecules, then -(q*lambdaCoef) for the removed ones.
 * firstCall: base = reference sums, else the current new sums.  scale = 1. */
int gomcb200_mol_exchange_reciprocal(gomcb200_engine *e, int box, int n,
                                     const double *w, const double *x,
                                     const double *y, const double *z,
                                     int firstCall, double scale,
                                     double *energyRecipNew);
/* CallChangeLambdaMolReciprocalGPU (CalculateEwaldCUDAKernel.cuh:37-43) ==
 * Ewald::ChangeLambdaRecip, src/Ewald.cpp:534-585;
 * lambdaCoef = sqrt(lambdaNew) - sqrt(lambdaOld). */
int gomcb200_change_lambda_mol_reciprocal(gomcb200_engine *e, int box,
                                          int molIndex, const double *x,
                                          const double *y, const double *z,
                                          double lambdaCoef,
                                          double *energyRecipNew);
/* Ewald::ChangeRecip, src/Ewald.cpp:589-642 (host-only in the reference):
 * energyRecip[s], s < nStates <= 64, of the resident molecule molIndex with its
 * charges scaled by sqrt(lambdaCoul[s]) - sqrt(lambdaCoul[iState]). */
int gomcb200_change_recip(gomcb200_engine *e, int box, int molIndex, int nStates,
                          const double *lambdaCoul, int iState,
                          double *energyRecip);
/* The lambda = 1 self and correction energies of the resident molecule molIndex that
 * Ewald::ChangeSelf (src/Ewald.cpp:1395-1417) and Ewald::ChangeCorrection (:1089-1122)
 * scale by (lambda_Coul[s] - lambda_Coul[iState]) for every free-energy state. */
int gomcb200_change_self_correction(gomcb200_engine *e, int box, int molIndex,
                                    double *enSelf, double *correction);
/* Ewald::SwapCorrection (src/Ewald.cpp:1311-1335 and :1340-1370; charges of
 * molIndex, trial coordinates x/y/z) and Ewald::SwapSelf (:1375-1391). */
int gomcb200_swap_correction(gomcb200_engine *e, int box, int molIndex,
                             const double *x, const double *y, const double *z,
                             double *correction, double *self);
/* One swap trial in one box: Swap{Dest,Source}Recip + SwapCorrection + SwapSelf
 * (the calls MoleculeTransfer::CalcEn makes per box,
 * src/moves/MoleculeTransfer.h:120-140) with a single synchronisation. */
int gomcb200_swap_trial(gomcb200_engine *e, int box, int molIndex,
                        const double *x, const double *y, const double *z,
                        int insert, double *energyRecipNew, double *correction,
                        double *self);
/* state machine, src/Ewald.cpp:1021-1053, :1420-1487 */
int gomcb200_set_recip_ref(gomcb200_engine *e, int box);    /* SetRecipRef / CopyCurrentToRefCUDA */
int gomcb200_copy_recip(gomcb200_engine *e, int box);       /* CopyRecip / CopyRefToNewCUDA       */
int gomcb200_update_recip(gomcb200_engine *e, int box);     /* UpdateRecip / UpdateRecipCUDA      */
int gomcb200_update_recip_vec(gomcb200_engine *e, int box); /* UpdateRecipVec / UpdateRecipVecCUDA */
/* Ewald::BoxSelf (src/Ewald.cpp:1125-1163) and the box sum of
 * Ewald::MolCorrection (src/Ewald.cpp:1056-1085). */
int gomcb200_box_self_correction(gomcb200_engine *e, int box, double *self,
                                 double *correction);

/* ---- literal drop-ins: host buffers in, scalars out -------------------- */
/* CallBoxInterGPU with the reference's argument meaning: x/y/z are the full
 * System::coordinates arrays (nAtoms each), axis the current box axes. */
int gomcb200_call_box_inter(gomcb200_engine *e, int box, const double *x,
                            const double *y, const double *z,
                            const double axis[3], double *REn, double *LJEn);
/* CallBoxReciprocalSumsGPU + BoxReciprocal */
int gomcb200_call_box_reciprocal_sums(gomcb200_engine *e, int box,
                                      const double *x, const double *y,
                                      const double *z, double *energyRecip);
/* CallBoxForceGPU: also downloads atom and molecule forces (may be NULL). */
int gomcb200_call_box_force(gomcb200_engine *e, int box, const double *x,
                            const double *y, const double *z,
                            const double axis[3], double *REn, double *LJEn,
                            double *aForcex, double *aForcey, double *aForcez,
                            double *mForcex, double *mForcey,
                            double *mForcez);
/* One full-box evaluation = BoxInter + BoxReciprocalSums + BoxReciprocal
 * (the E1 metric of SURVEY.md section 8d) with host coordinates in. */
int gomcb200_call_full_box_energy(gomcb200_engine *e, int box,
                                  const double *x, const double *y,
                                  const double *z, double *LJEn, double *REn,
                                  double *energyRecip);

/* ---- literal drop-ins of the reciprocal seam ------------------------------
 * The reference's Call*GPU functions of src/GPU/CalculateEwaldCUDAKernel.cuh take the
 * k list, the box's point charges and the moved molecule as HOST arrays and return the
 * sums to HOST arrays; these entry points keep exactly that convention, so the shim of
 * integration/ (the reference's own symbols on top of this library) needs no knowledge of
 * the topology.  sumRnew / sumInew may be NULL (no device -> host copy). */
/* The k list Ewald::RecipInit built on the host (kx[box] ... prefact[box], imageSize[box])
 * into the NEW k set: what CallBoxReciprocalSetupGPU receives (CalculateEwaldCUDAKernel.cuh:
 * 27-33).  The integer (a, b, c) structure is recovered from the list; the prefactors are
 * the host's. */
int gomcb200_set_kvectors(gomcb200_engine *e, int box, int n, const double *kx,
                          const double *ky, const double *kz, const double *hsqr,
                          const double *prefact);
/* CallBoxReciprocalSetupGPU (newSet = 1) / CallBoxReciprocalSumsGPU (newSet = 0) with the
 * reference's arguments: n point charges q (already times lambdaCoef) at x, y, z. */
int gomcb200_call_box_reciprocal_points(gomcb200_engine *e, int box, int newSet, int n,
                                        const double *x, const double *y, const double *z,
                                        const double *q, double *sumRnew, double *sumInew,
                                        double *energyRecip);
/* CallMolReciprocalGPU (CalculateEwaldCUDAKernel.cuh:40-44): old and new coordinates and
 * the charges (times lambdaCoef) of the moved molecule. */
int gomcb200_call_mol_reciprocal(gomcb200_engine *e, int box, int len, const double *q,
                                 const double *oldX, const double *oldY, const double *oldZ,
                                 const double *newX, const double *newY, const double *newZ,
                                 double *sumRnew, double *sumInew, double *energyRecipNew);
/* CallSwapReciprocalGPU (CalculateEwaldCUDAKernel.cuh:54-57). */
int gomcb200_call_swap_reciprocal(gomcb200_engine *e, int box, int len, const double *q,
                                  const double *x, const double *y, const double *z,
                                  int insert, double *sumRnew, double *sumInew,
                                  double *energyRecipNew);
/* CallMolExchangeReciprocalGPU (CalculateEwaldCUDAKernel.cuh:45-48): host sums -> device. */
int gomcb200_set_recip_sums(gomcb200_engine *e, int box, int which, const double *sumR,
                            const double *sumI, int n);
/* Host forces / torques -> the resident buffers (the molForceRef / molTorqueRef arguments of
 * CallTranslateParticlesGPU / CallRotateParticlesGPU, TransformParticlesCUDAKernel.cuh:20-37). */
int gomcb200_set_forces(gomcb200_engine *e, int which, const double *x, const double *y,
                        const double *z, int first, int count);

/* ---- multi-GPU sharding (one engine per GPU, one process per GPU) -------- */
/* Rank `rank` of `world` evaluates its share of every full-box sweep:
 * a contiguous slab of cells for the pair path and a contiguous block of
 * (kx,ky) rows of the k list for the structure factor.  Coordinates are
 * replicated; energies returned by box_inter / box_force /
 * box_reciprocal_sums / call_full_box_energy are then PARTIAL sums that the
 * caller all-reduces (3 doubles, NCCL).  sumRnew/sumInew stay distributed:
 * entries of k-vectors owned by other ranks are zero.  Entry points that need the
 * complete sums or every atom's force (molecule/swap trials, the reciprocal deltas,
 * box_force, box_force_reciprocal, virial_reciprocal, calculate_torque, mp/bm_coeff)
 * return GOMCB200_EINVAL while world > 1.  world == 1 restores the single-GPU
 * behaviour. */
int gomcb200_set_shard(gomcb200_engine *e, int rank, int world);
/* The collective inside the engine.  gomcb200_comm_unique_id fills 128 bytes (an NCCL unique
 * id; call it on one rank and hand the bytes to the others by any means -- MPI, a socket,
 * shared memory, or plain memory between the threads of one process).  gomcb200_set_comm
 * makes the engine rank `rank` of a `world`-rank NCCL communicator (libnccl.so.2, bound at run
 * time) and implies gomcb200_set_shard(rank, world).  From then on box_inter,
 * box_reciprocal_setup/_sums and call_full_box_energy all-reduce their energies on the
 * engine's stream before the one device-to-host copy, so every rank returns the COMPLETE
 * values, and the structure factor runs as the slab-sharded non-uniform FFT (the pruned
 * slabs are all-gathered; every rank ends with the complete sumRnew/sumInew).  The
 * MultiParticle path runs sharded too (single-box engines): box_force computes the forces of
 * this rank's cell slab and all-reduces the atom force arrays, box_force_reciprocal /
 * calculate_torque / mp_transform / mp_coeff then run replicated on the complete data, so
 * every rank takes the same accept/reject decision without exchanging coordinates.  One engine
 * per GPU, one host thread per engine: the ranks may be processes (torchrun, MPI) or the
 * threads of ONE process -- GOMC is a single process (src/Main.cpp:318-326) and reaches all
 * GPUs of a box this way (gomc_b200/host/multi_gpu_test.cpp).  world == 1 drops the
 * communicator. */
int gomcb200_comm_unique_id(void *id128);
int gomcb200_set_comm(gomcb200_engine *e, const void *id128, int rank, int world);
/* Tell the engine that resident coordinates are to be treated as changed
 * (forces cell re-binning and re-packing on the next sweep), as after a
 * device-side MultiParticle transform. */
int gomcb200_mark_coords_changed(gomcb200_engine *e);

/* ---- tuning / introspection (tests and bench only) ---------------------- */
/* algorithm for the structure-factor build (the sums behind BoxReciprocalSetup/
 * BoxReciprocalSums, src/Ewald.cpp:213-330) and the reciprocal force:
 *   4 = default: 5 where it applies (orthogonal box), else 2, else 0 (triclinic cell);
 *   5 = non-uniform FFT (spread on the FP64 tensor path + pruned FFT; agrees with the direct
 *       sums to ~1e-13 of max |S|, energy ~1e-15 relative);
 *   0 = direct sincos per (atom,k) (the reference's algorithm);
 *   1 = factorised per-axis phases on the FP64 CUDA cores;
 *   2 = the same factorisation on the FP64 MMA path (DMMA);
 *   3 = byte-sliced fixed point on the INT8 tensor cores (tcgen05 + TMEM; agrees with 2 to
 *       < 1e-11 relative). */
int gomcb200_set_recip_algo(gomcb200_engine *e, int algo);
/* pair-sweep kernel of BoxInter / BoxForce / VirialCalc on orthogonal boxes: 1 = default,
 * k_pair_box2 (TMA-staged cells, FP32 candidate filter, tabulated Ewald real-space terms);
 * 0 = the first kernel (FP64 candidate tests, library erfc), always used for triclinic
 * boxes.  Both decide InRcut on the same FP64 r^2; energies agree to ~1e-13 relative. */
int gomcb200_set_pair_algo(gomcb200_engine *e, int algo);
/* work threshold (charged atoms x k-vectors) of algorithm 4 */
int gomcb200_set_recip_auto_work(gomcb200_engine *e, double work);
/* CUDA-event time (ms) of the kernels launched by the last call, and the
 * device time of its dominant kernel. */
int gomcb200_last_timing(const gomcb200_engine *e, float *totalMs,
                         float *dominantMs);
/* enable/disable per-call event timing (off by default) */
int gomcb200_enable_timing(gomcb200_engine *e, int on);

#ifdef __cplusplus
}
#endif
#endif /* GOMC_B200_H */
