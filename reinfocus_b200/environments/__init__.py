"""Drop-in counterparts of reinfocus.environments: the Gymnasium env / vector env and the
six strategy objects they compose. Only FocusObserver touches the GPU (through
FastRenderer); the rest is the same small NumPy glue as in the reference."""
