"""Drop-in mirror of the reference's ``spotify/`` hot path (SpotifyModel + train_step)."""
