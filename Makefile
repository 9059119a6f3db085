# Same entry points as the reference's Makefile (Makefile:21-32 there): `make` builds the solver,
# `make check` compares ./av_vels.dat and ./final_state.dat with the reference results
# (REF_AV_VELS_FILE / REF_FINAL_STATE_FILE / AV_VELS_FILE / FINAL_STATE_FILE as in the reference).
# The product lives in hpc-lattice-boltzmann_b200/; `make oracle` builds the CPU checkers.

PKG = hpc-lattice-boltzmann_b200

all:
	$(MAKE) -C $(PKG)
	ln -sf $(PKG)/d2q9-bgk.exe d2q9-bgk.exe
	ln -sf $(PKG)/d2q9-bgk.exe d2q9-bgk

check:
	$(MAKE) -C $(PKG) check FINAL_STATE_FILE=$(abspath $(or $(FINAL_STATE_FILE),./final_state.dat)) \
	    AV_VELS_FILE=$(abspath $(or $(AV_VELS_FILE),./av_vels.dat)) \
	    $(if $(REF_FINAL_STATE_FILE),REF_FINAL_STATE_FILE=$(abspath $(REF_FINAL_STATE_FILE))) \
	    $(if $(REF_AV_VELS_FILE),REF_AV_VELS_FILE=$(abspath $(REF_AV_VELS_FILE)))

oracle:
	$(MAKE) -C oracle
	$(MAKE) -C oracle ref

clean:
	$(MAKE) -C $(PKG) clean
	rm -f d2q9-bgk.exe d2q9-bgk

.PHONY: all check oracle clean
