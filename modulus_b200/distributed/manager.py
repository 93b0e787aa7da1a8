"""DistributedManager -- process bookkeeping for the partitioned-graph path.

A compact equivalent of `physicsnemo.distributed.DistributedManager`
(reference: physicsnemo/distributed/manager.py:37-775) with the members the
MeshGraphNet / DistributedGraph path touches: `initialize()`, `is_initialized()`, `rank`,
`world_size`, `local_rank`, `device`, `group()/group_size()/group_rank()`,
`create_process_subgroup()`, `cleanup()`.  One process per GPU, NCCL on CUDA, gloo on CPU
(the reference's rule, manager.py:292-298).  Pure host logic, no arithmetic.
"""
from __future__ import annotations

import atexit
import os
import warnings
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class PhysicsNeMoUndefinedGroupError(Exception):
    def __init__(self, name: str):
        super().__init__(f"Cannot query process group '{name}' before it is explicitly created.")


class PhysicsNeMoUninitializedDistributedManagerWarning(Warning):
    def __init__(self):
        super().__init__(
            "A DistributedManager object is being instantiated before this singleton class has been "
            "initialized. Call DistributedManager.initialize() first."
        )


class DistributedManager:
    """Borg singleton holding rank / device / named process groups."""

    _shared_state: Dict = {}

    def __new__(cls):
        obj = super().__new__(cls)
        obj.__dict__ = cls._shared_state
        if not hasattr(obj, "_rank"):
            obj._rank = 0
            obj._world_size = 1
            obj._local_rank = 0
            obj._distributed = False
            obj._device = torch.device(f"cuda:0" if torch.cuda.is_available() else "cpu")
            obj._cuda = torch.cuda.is_available()
            obj._groups = {}
            obj._group_ranks = {}
            obj._is_initialized = False
            obj._owns_pg = False
        return obj

    def __init__(self):
        if not self._is_initialized:
            warnings.warn(PhysicsNeMoUninitializedDistributedManagerWarning().args[0])

    # ------------------------------------------------------------------ properties
    @property
    def rank(self) -> int:
        return self._rank

    @property
    def local_rank(self) -> int:
        return self._local_rank

    @property
    def world_size(self) -> int:
        return self._world_size

    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def distributed(self) -> bool:
        return self._distributed

    @property
    def cuda(self) -> bool:
        return self._cuda

    @property
    def group_names(self) -> List[str]:
        return list(self._groups.keys())

    @classmethod
    def is_initialized(cls) -> bool:
        return cls._shared_state.get("_is_initialized", False)

    # ------------------------------------------------------------------ groups
    def group(self, name: Optional[str] = None):
        """Process group by name; None is the default (world) group (manager.py:219-232)."""
        if name is None:
            return None
        if name in self._groups:
            return self._groups[name]
        raise PhysicsNeMoUndefinedGroupError(name)

    def group_size(self, name: Optional[str] = None) -> int:
        if name is None:
            return self._world_size
        return dist.get_world_size(group=self.group(name))

    def group_rank(self, name: Optional[str] = None) -> int:
        if name is None:
            return self._rank
        return dist.get_rank(group=self.group(name))

    def group_name(self, group=None) -> Optional[str]:
        for k, v in self._groups.items():
            if v is group:
                return k
        return None

    def create_process_subgroup(self, name: str, size: int, group_name: Optional[str] = None,
                                verbose: bool = False):
        """Split the parent group into consecutive sub-groups of `size` ranks (manager.py:570-644)."""
        if name in self._groups:
            raise AssertionError(f"Group with name {name} already exists")
        parent = self.group(group_name)
        parent_size = self.group_size(group_name)
        if parent_size % size != 0:
            raise AssertionError(f"Cannot divide group size {parent_size} evenly into subgroups of size {size}")
        if not self._distributed:
            self._groups[name] = None
            self._group_ranks[name] = [[0]]
            return
        parent_ranks = dist.get_process_group_ranks(parent if parent is not None else dist.group.WORLD)
        self._group_ranks[name] = []
        for i in range(parent_size // size):
            ranks = parent_ranks[i * size:(i + 1) * size]
            pg = dist.new_group(ranks=ranks)  # every rank must take part in every new_group call
            self._group_ranks[name].append(ranks)
            if self._rank in ranks:
                self._groups[name] = pg
        if verbose and self._rank == 0:
            print(f"Process group '{name}': {self._group_ranks[name]}")

    # ------------------------------------------------------------------ init / teardown
    @staticmethod
    def initialize():
        """Initialise from torchrun-style environment variables (manager.py:363-427; the
        SLURM / OpenMPI fallbacks of the reference are host plumbing outside this path)."""
        if DistributedManager.is_initialized():
            warnings.warn("Distributed manager is already intialized")
            return
        rank = int(os.environ.get("RANK", os.environ.get("OMPI_COMM_WORLD_RANK", 0)))
        world_size = int(os.environ.get("WORLD_SIZE", os.environ.get("OMPI_COMM_WORLD_SIZE", 1)))
        local_rank = int(os.environ.get("LOCAL_RANK", os.environ.get("OMPI_COMM_WORLD_LOCAL_RANK", 0)))
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = os.environ.get("MASTER_PORT", "12355")
        DistributedManager.setup(rank, world_size, local_rank, addr, port)

    @staticmethod
    def setup(rank=0, world_size=1, local_rank=None, addr="127.0.0.1", port="12355", backend=None):
        os.environ["MASTER_ADDR"] = addr
        os.environ["MASTER_PORT"] = str(port)
        DistributedManager._shared_state["_is_initialized"] = True
        m = DistributedManager()
        m._is_initialized = True  # (the first instantiation fills in defaults, this flag among them)
        m._rank, m._world_size = rank, world_size
        m._local_rank = local_rank if local_rank is not None else (
            rank % torch.cuda.device_count() if torch.cuda.is_available() else 0)
        m._distributed = world_size > 1 and dist.is_available()
        m._cuda = torch.cuda.is_available()
        m._device = torch.device(f"cuda:{m._local_rank}" if m._cuda else "cpu")
        if m._cuda:
            torch.cuda.set_device(m._device)
        if m._distributed and not dist.is_initialized():
            backend = backend or ("nccl" if m._cuda else "gloo")
            kwargs = {"device_id": m._device} if (m._cuda and backend == "nccl") else {}
            dist.init_process_group(backend, rank=rank, world_size=world_size, **kwargs)
            m._owns_pg = True
        atexit.register(DistributedManager.cleanup)

    @staticmethod
    def cleanup():
        st = DistributedManager._shared_state
        if st.get("_is_initialized", False):
            if st.get("_owns_pg", False) and dist.is_available() and dist.is_initialized():
                try:
                    dist.destroy_process_group()
                except Exception:
                    pass
            st.clear()
