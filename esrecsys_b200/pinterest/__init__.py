"""Drop-in mirror of the reference's ``pinterest/`` hot path (STLModel scoring + train_step loss)."""
