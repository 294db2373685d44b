/* The reference reaches into PETSc's private DM_DA to build element connectivity
 * (LinearElasticity.cc:785-839); the B200 classes index the structured grid directly, so nothing is needed here. */
