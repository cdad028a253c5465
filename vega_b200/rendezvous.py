"""Tracker-style TCP rendezvous for one-process-per-GPU executors.

In the reference the master's MapOutputTracker listens on a TCP socket and every executor asks it
where a shuffle's map outputs live (src/map_output_tracker.rs:68-93 client, :95-166 server: connect
in a retry loop, one request, one reply, the server polls with a 1 ms sleep until the answer exists).
With the exchange inside libvega_b200 the only thing the executors must agree on before the first
shuffle is the 128-byte NCCL unique id — so the tracker channel carries exactly that:

    master (rank 0)  : id = Context.comm_unique_id();  TrackerServer(addr, world, id).serve()
    every rank       : id = fetch_unique_id(addr, rank, world);  Context.comm_init(rank, world, id)

Wire format (little endian): request = magic u32 | rank u32 | world u32; reply = magic u32 | n u32 | n bytes.
The server answers every rank once (rank 0 may ask too) and stops after `world` distinct ranks, or on close().
No torch, no CUDA: plain sockets, testable on CPU.
"""
import socket
import struct
import threading
import time

MAGIC = 0x76423230          # "vB20"
_REQ = struct.Struct("<III")
_REP = struct.Struct("<II")


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed the tracker connection")
        buf += chunk
    return bytes(buf)


class TrackerServer:
    """The master side: hands `payload` (the NCCL unique id) to each of `world` ranks exactly as
    MapOutputTracker::server hands out URI lists — one request/reply per connection."""

    def __init__(self, addr, world, payload, timeout=120.0):
        self.world, self.payload, self.timeout = world, bytes(payload), timeout
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind(addr)
        self.sock.listen(world + 8)
        self.addr = self.sock.getsockname()
        self.seen = set()
        self.error = None
        self._thread = None

    def serve(self, background=True):
        if background:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
            return self
        self._run()
        return self

    def _run(self):
        deadline = time.time() + self.timeout
        try:
            while len(self.seen) < self.world:
                self.sock.settimeout(max(0.05, deadline - time.time()))
                try:
                    conn, _ = self.sock.accept()
                except socket.timeout:
                    self.error = TimeoutError(f"tracker: only ranks {sorted(self.seen)} of {self.world} showed up")
                    return
                with conn:
                    conn.settimeout(10.0)
                    try:
                        magic, rank, world = _REQ.unpack(_recv_exact(conn, _REQ.size))
                    except Exception:
                        continue                      # stray connection: ignore, like a failed capnp read
                    if magic != MAGIC or world != self.world or rank >= self.world:
                        conn.sendall(_REP.pack(MAGIC, 0))
                        continue
                    conn.sendall(_REP.pack(MAGIC, len(self.payload)) + self.payload)
                    self.seen.add(rank)
        finally:
            self.sock.close()

    def join(self, timeout=None):
        if self._thread:
            self._thread.join(timeout)
        if self.error:
            raise self.error

    def close(self):
        try:
            self.sock.close()
        except Exception:
            pass


def fetch_unique_id(addr, rank, world, timeout=120.0):
    """The executor side (MapOutputTracker::client): connect in a retry loop until the master listens,
    send (rank, world), read the id."""
    deadline = time.time() + timeout
    last = None
    while time.time() < deadline:
        try:
            with socket.create_connection(addr, timeout=5.0) as s:
                s.sendall(_REQ.pack(MAGIC, rank, world))
                magic, n = _REP.unpack(_recv_exact(s, _REP.size))
                if magic != MAGIC:
                    raise ConnectionError("not a vega_b200 tracker")
                if n == 0:
                    raise ValueError(f"tracker refused rank {rank} of world {world}")
                return _recv_exact(s, n)
        except (ConnectionRefusedError, ConnectionResetError, socket.timeout, OSError) as e:
            last = e
            time.sleep(0.02)
    raise TimeoutError(f"no tracker at {addr}: {last}")


def bootstrap(sc, rank, world, master_addr):
    """One call per process: rank 0 creates the NCCL id and serves it, everyone fetches it and joins the
    communicator of `sc` (vb_ctx_comm_init).  Replaces the torch.distributed broadcast in Context.comm_init."""
    server = None
    if rank == 0:
        server = TrackerServer(master_addr, world, sc.comm_unique_id()).serve()
    uid = fetch_unique_id(master_addr, rank, world)
    sc.comm_init(rank, world, unique_id=uid)
    if server:
        server.join(30.0)
    return uid
