"""Drop-in counterparts of reinfocus.graphics for the hot path."""
