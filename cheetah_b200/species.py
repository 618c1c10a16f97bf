"""Particle species (mirror of cheetah/particles/species.py:12-149, hot-path subset)."""

from __future__ import annotations

import torch
from torch import nn

# CODATA 2022, the values scipy.constants hands to the reference (species.py:5-9)
ELECTRON_MASS_EV = 510998.95069
PROTON_MASS_EV = 938272089.4300001
DEUTERON_MASS_EV = 1875612945.0
ELEMENTARY_CHARGE = 1.602176634e-19
EV_TO_KG = 1.7826619216278975e-36
SPEED_OF_LIGHT = 299792458.0
EPSILON_0 = 8.8541878188e-12


class Species(nn.Module):
    """Named particle species defined by charge (in e) and mass (in eV)."""

    known = {
        "electron": {"num_elementary_charges": -1, "mass_eV": ELECTRON_MASS_EV},
        "positron": {"num_elementary_charges": 1, "mass_eV": ELECTRON_MASS_EV},
        "proton": {"num_elementary_charges": 1, "mass_eV": PROTON_MASS_EV},
        "antiproton": {"num_elementary_charges": -1, "mass_eV": PROTON_MASS_EV},
        "deuteron": {"num_elementary_charges": 1, "mass_eV": DEUTERON_MASS_EV},
    }

    def __init__(
        self,
        name: str,
        num_elementary_charges: torch.Tensor | None = None,
        charge_coulomb: torch.Tensor | None = None,
        mass_eV: torch.Tensor | None = None,
        mass_kg: torch.Tensor | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        super().__init__()
        factory_kwargs = {"device": device, "dtype": dtype}
        self.name = name
        if name in self.known:
            assert all(
                v is None for v in (num_elementary_charges, charge_coulomb, mass_eV, mass_kg)
            ), "Known particle species should not have charge and mass provided."
            charges = torch.tensor(self.known[name]["num_elementary_charges"], **factory_kwargs)
            mass = torch.tensor(self.known[name]["mass_eV"], **factory_kwargs)
        else:
            assert (num_elementary_charges is not None) != (
                charge_coulomb is not None
            ), "Provide exactly one of num_elementary_charges and charge_coulomb."
            assert (mass_eV is not None) != (
                mass_kg is not None
            ), "Provide exactly one of mass_eV and mass_kg."
            charges = (
                num_elementary_charges
                if num_elementary_charges is not None
                else charge_coulomb / ELEMENTARY_CHARGE
            )
            mass = mass_eV if mass_eV is not None else mass_kg / EV_TO_KG
        self.register_buffer("num_elementary_charges", charges)
        self.register_buffer("mass_eV", mass)

    @property
    def mass_kg(self) -> torch.Tensor:
        return self.mass_eV * EV_TO_KG

    @property
    def charge_coulomb(self) -> torch.Tensor:
        return self.num_elementary_charges * ELEMENTARY_CHARGE

    def clone(self) -> "Species":
        """New ``Species`` object for an outgoing beam (element.py:190).  The two scalar tensors
        are SHARED with the original instead of copied: nothing in this package modifies them in
        place, and two device-side copies per tracked section are pure launch overhead."""
        copy = self.__class__.__new__(self.__class__)
        nn.Module.__init__(copy)
        object.__setattr__(copy, "name", self.name)
        copy._buffers["num_elementary_charges"] = self._buffers["num_elementary_charges"]
        copy._buffers["mass_eV"] = self._buffers["mass_eV"]
        return copy

    def __repr__(self) -> str:
        return (
            f"Species(name={self.name!r}, num_elementary_charges="
            f"{self.num_elementary_charges!r}, mass_eV={self.mass_eV!r})"
        )
