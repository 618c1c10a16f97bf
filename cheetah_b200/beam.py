"""Beam state containers (mirror of cheetah/particles/{particle_beam,parameter_beam}.py).

Only what the ``Segment.track`` hot path needs is mirrored: the constructor contract
(particle_beam.py:60-106, parameter_beam.py ctor), the coordinate views, the
survival-weighted first/second moments (particle_beam.py:1699-1805) and two set-up
generators.  Generators, plotting and file I/O of the reference stay the reference's job
(SURVEY.md 2, rows 8 and 10).
"""

from __future__ import annotations

import torch
from torch import nn

from .species import Species


class Beam(nn.Module):
    @property
    def relativistic_gamma(self) -> torch.Tensor:
        return self.energy / self.species.mass_eV

    @property
    def relativistic_beta(self) -> torch.Tensor:
        # cheetah/particles/beam.py:328-336
        gamma = self.relativistic_gamma
        beta = torch.ones_like(gamma)
        nonzero = gamma.abs() > 0
        beta[nonzero] = (1 - gamma[gamma > 0].square().reciprocal()).sqrt()
        return beta

    @property
    def p0c(self) -> torch.Tensor:
        return self.relativistic_beta * self.relativistic_gamma * self.species.mass_eV


class ParticleBeam(Beam):
    """Beam of macroparticles: ``particles (..., N, 7)`` rows ``[x, px, y, py, tau, delta, 1]``."""

    def __init__(
        self,
        particles: torch.Tensor,
        energy: torch.Tensor,
        particle_charges: torch.Tensor | None = None,
        survival_probabilities: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        super().__init__()
        assert (
            particles.shape[-2] > 0 and particles.shape[-1] == 7
        ), "Particle vectors must be 7-dimensional."
        device = device if device is not None else particles.device
        dtype = dtype if dtype is not None else particles.dtype
        factory_kwargs = {"device": device, "dtype": dtype}
        self.species = species if species is not None else Species("electron", **factory_kwargs)
        self.register_buffer("particles", particles)
        self.register_buffer("energy", energy)
        self.register_buffer(
            "particle_charges",
            particle_charges
            if particle_charges is not None
            else torch.full(
                (particles.shape[-2],), float(self.species.charge_coulomb), **factory_kwargs
            ),
        )
        self.register_buffer(
            "survival_probabilities",
            survival_probabilities
            if survival_probabilities is not None
            else torch.ones(particles.shape[-2], **factory_kwargs),
        )
        self.register_buffer("s", s if s is not None else torch.tensor(0.0, **factory_kwargs))
        # None = unknown; the tracker checks once whether particles[..., 6] == 1 and caches it
        self._unit_seventh: bool | None = None

    @classmethod
    def _from_tracking(cls, particles, energy, particle_charges, survival_probabilities, s,
                       species, unit_seventh=None) -> "ParticleBeam":
        """Constructor used by the tracker for outgoing beams: same object as ``__init__`` builds,
        without its argument checks and ``register_buffer`` bookkeeping (~60 us per call)."""
        beam = cls.__new__(cls)
        nn.Module.__init__(beam)
        beam._modules["species"] = species
        buffers = beam._buffers
        buffers["particles"] = particles
        buffers["energy"] = energy
        buffers["particle_charges"] = particle_charges
        buffers["survival_probabilities"] = survival_probabilities
        buffers["s"] = s
        object.__setattr__(beam, "_unit_seventh", unit_seventh)
        return beam

    # ---- set-up helpers (host-side convenience, not part of the accelerated path) --------
    @classmethod
    def from_distribution(
        cls,
        mu: torch.Tensor,
        cov: torch.Tensor,
        num_particles: int = 100_000,
        energy: torch.Tensor | None = None,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
        generator: torch.Generator | None = None,
    ) -> "ParticleBeam":
        """Gaussian beam with mean ``mu (6,)`` and covariance ``cov (6, 6)``.

        Unlike the reference (particle_beam.py:357-431) the sample moments are not matched
        to (mu, cov) exactly; samples are drawn through a Cholesky factor.
        """
        dtype = dtype if dtype is not None else torch.get_default_dtype()
        factory_kwargs = {"device": device, "dtype": dtype}
        species = species if species is not None else Species("electron", **factory_kwargs)
        energy = energy if energy is not None else torch.tensor(1e8, **factory_kwargs)
        if total_charge is None:
            total_charge = species.charge_coulomb * num_particles
        total_charge = torch.as_tensor(total_charge, **factory_kwargs)
        particle_charges = (
            torch.ones((*total_charge.shape, num_particles), **factory_kwargs)
            * total_charge.unsqueeze(-1)
            / num_particles
        )
        factor = torch.linalg.cholesky(
            cov.to(torch.float64).cpu() + 1e-300 * torch.eye(6, dtype=torch.float64)
        )
        standard = torch.randn(num_particles, 6, dtype=torch.float64, generator=generator)
        samples = standard @ factor.mT + mu.to(torch.float64).cpu()
        particles = torch.cat([samples, torch.ones(num_particles, 1, dtype=torch.float64)], dim=-1)
        beam = cls(
            particles.to(**factory_kwargs),
            energy.to(**factory_kwargs),
            particle_charges=particle_charges,
            s=s,
            species=species,
            **factory_kwargs,
        )
        beam._unit_seventh = True
        return beam

    @classmethod
    def from_parameters(
        cls,
        num_particles: int = 100_000,
        mu_x=0.0, mu_px=0.0, mu_y=0.0, mu_py=0.0, mu_tau=0.0, mu_p=0.0,
        sigma_x=175e-6, sigma_px=4e-6, sigma_y=175e-6, sigma_py=4e-6,
        sigma_tau=8e-6, sigma_p=2e-3,
        cov_xpx=0.0, cov_ypy=0.0, cov_taup=0.0,
        energy: torch.Tensor | None = None,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
        generator: torch.Generator | None = None,
    ) -> "ParticleBeam":
        """Defaults follow particle_beam.py:199-216."""
        f = lambda v: float(v)  # noqa: E731
        mu = torch.tensor([f(mu_x), f(mu_px), f(mu_y), f(mu_py), f(mu_tau), f(mu_p)], dtype=torch.float64)
        cov = torch.zeros(6, 6, dtype=torch.float64)
        for i, sigma in enumerate((sigma_x, sigma_px, sigma_y, sigma_py, sigma_tau, sigma_p)):
            cov[i, i] = f(sigma) ** 2
        cov[0, 1] = cov[1, 0] = f(cov_xpx)
        cov[2, 3] = cov[3, 2] = f(cov_ypy)
        cov[4, 5] = cov[5, 4] = f(cov_taup)
        return cls.from_distribution(
            mu, cov, num_particles, energy, total_charge, s, species, device, dtype, generator
        )

    @classmethod
    def from_twiss(
        cls,
        num_particles: int = 100_000,
        beta_x=0.0, alpha_x=0.0, emittance_x=7.1971891e-13,
        beta_y=0.0, alpha_y=0.0, emittance_y=7.1971891e-13,
        sigma_tau=1e-6, sigma_p=1e-6, cov_taup=0.0,
        energy: torch.Tensor | None = None,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
        generator: torch.Generator | None = None,
    ) -> "ParticleBeam":
        """Twiss -> second moments as in particle_beam.py:499-533 (no dispersion terms)."""
        f = lambda v: float(v)  # noqa: E731
        return cls.from_parameters(
            num_particles,
            sigma_x=(f(beta_x) * f(emittance_x)) ** 0.5,
            sigma_px=(f(emittance_x) * (1 + f(alpha_x) ** 2) / f(beta_x)) ** 0.5,
            sigma_y=(f(beta_y) * f(emittance_y)) ** 0.5,
            sigma_py=(f(emittance_y) * (1 + f(alpha_y) ** 2) / f(beta_y)) ** 0.5,
            sigma_tau=sigma_tau, sigma_p=sigma_p,
            cov_xpx=-f(emittance_x) * f(alpha_x),
            cov_ypy=-f(emittance_y) * f(alpha_y),
            cov_taup=cov_taup,
            energy=energy, total_charge=total_charge, s=s, species=species,
            device=device, dtype=dtype, generator=generator,
        )

    # ---- views and moments -----------------------------------------------------------------
    @property
    def num_particles(self) -> int:
        return self.particles.shape[-2]

    def __len__(self) -> int:
        return int(self.num_particles)

    @property
    def total_charge(self) -> torch.Tensor:
        return (self.particle_charges * self.survival_probabilities).sum(dim=-1)

    @property
    def num_particles_survived(self) -> torch.Tensor:
        return self.survival_probabilities.sum(dim=-1)

    def _coordinate(self, index: int) -> torch.Tensor:
        return self.particles[..., index]

    def _mean(self, index: int) -> torch.Tensor:
        w = self.survival_probabilities
        return (self._coordinate(index) * w).sum(dim=-1) / w.sum(dim=-1)

    def _std(self, index: int) -> torch.Tensor:
        # unbiased survival-weighted std: cheetah/utils/statistics.py:30-62
        w = self.survival_probabilities
        v = self._coordinate(index)
        sum_w = w.sum(dim=-1)
        mean = (v * w).sum(dim=-1) / sum_w
        correction = sum_w - w.square().sum(dim=-1) / sum_w
        return ((w * (v - mean.unsqueeze(-1)).square()).sum(dim=-1) / correction).sqrt()

    def clone(self) -> "ParticleBeam":
        beam = self.__class__(
            particles=self.particles.clone(),
            energy=self.energy.clone(),
            particle_charges=self.particle_charges.clone(),
            survival_probabilities=self.survival_probabilities.clone(),
            s=self.s.clone(),
            species=self.species.clone(),
        )
        beam._unit_seventh = self._unit_seventh
        return beam

    def __repr__(self) -> str:
        return (
            f"{self.__class__.__name__}(particles={tuple(self.particles.shape)}, "
            f"energy={self.energy!r}, s={self.s!r}, species={self.species!r})"
        )


for _i, _name in enumerate(("x", "px", "y", "py", "tau", "p")):
    setattr(ParticleBeam, _name, property(lambda self, i=_i: self._coordinate(i)))
    setattr(ParticleBeam, f"mu_{_name}", property(lambda self, i=_i: self._mean(i)))
    setattr(ParticleBeam, f"sigma_{_name}", property(lambda self, i=_i: self._std(i)))


class ParameterBeam(Beam):
    """Gaussian-moment beam: ``mu (..., 7)``, ``cov (..., 7, 7)`` (parameter_beam.py:8-60)."""

    def __init__(
        self,
        mu: torch.Tensor,
        cov: torch.Tensor,
        energy: torch.Tensor,
        total_charge: torch.Tensor | None = None,
        s: torch.Tensor | None = None,
        species: Species | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        super().__init__()
        device = device if device is not None else mu.device
        dtype = dtype if dtype is not None else mu.dtype
        factory_kwargs = {"device": device, "dtype": dtype}
        self.species = species if species is not None else Species("electron", **factory_kwargs)
        self.register_buffer("mu", mu)
        self.register_buffer("cov", cov)
        self.register_buffer("energy", energy)
        self.register_buffer(
            "total_charge",
            total_charge if total_charge is not None else torch.tensor(0.0, **factory_kwargs),
        )
        self.register_buffer("s", s if s is not None else torch.tensor(0.0, **factory_kwargs))

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(mu={self.mu!r}, energy={self.energy!r}, s={self.s!r})"
