"""Host mirrors of the reference's loss pieces that sit directly on the hot path's outputs
(SURVEY 8f): the MMA latency loss, the SSNT lattice loss and the CTC best alignment."""
