"""mpi4py stand-in over the torchrun environment (reference call sites: mm_diffusion/dist_util.py:26-47,
multimodal_datasets.py:48-87): COMM_WORLD.Get_rank / Get_size / rank / size / bcast / barrier.

bcast(obj, root) works for any world size launched by torchrun: rank `root` publishes the pickled object in a
torch.distributed.TCPStore on MASTER_ADDR:(MASTER_PORT + 29), the other ranks fetch it."""
import os
import pickle


class _Comm:
    def __init__(self):
        self._store = None
        self._seq = 0

    def Get_rank(self):
        return int(os.environ.get("RANK", os.environ.get("OMPI_COMM_WORLD_RANK", "0")))

    def Get_size(self):
        return int(os.environ.get("WORLD_SIZE", os.environ.get("OMPI_COMM_WORLD_SIZE", "1")))

    rank = property(Get_rank)
    size = property(Get_size)

    def _get_store(self):
        if self._store is None:
            from datetime import timedelta
            import torch.distributed as dist
            addr = os.environ.get("MMD_SHIM_ADDR", os.environ.get("MASTER_ADDR", "127.0.0.1"))
            port = int(os.environ.get("MMD_SHIM_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 29))
            self._store = dist.TCPStore(addr, port, self.Get_size(), is_master=(self.Get_rank() == 0),
                                        timeout=timedelta(seconds=120))
        return self._store

    def bcast(self, obj, root=0):
        if self.Get_size() == 1:
            return obj
        store = self._get_store()
        key = f"bcast{self._seq}"
        self._seq += 1
        if self.Get_rank() == root:
            store.set(key, pickle.dumps(obj))
            return obj
        return pickle.loads(store.get(key))

    def Barrier(self):
        self.bcast(None, 0)

    barrier = Barrier


class _MPI:
    COMM_WORLD = _Comm()


MPI = _MPI()
