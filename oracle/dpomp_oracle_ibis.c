/* placeholder: outer layers (SMC^2, MBP-IBIS) are added below */
