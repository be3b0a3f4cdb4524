"""External LOG endpoint of a running experiment (reference: the ``LOG`` message of maggy/core/rpc.py:490-502, which
the Hopsworks Jupyter front-end polls after the driver registered ``{hostIp, port, appId, secret}`` over REST,
maggy/core/environment/hopsworks.py:136-190).

Workers talk to the driver through shared memory, so the only socket this framework opens is this one, and only on
request (``MAGGY_LOG_SERVER=1`` / ``=host:port``, or automatically under a Hopsworks environment). Wire format:
4-byte big-endian length + UTF-8 JSON (no pickles cross this boundary). Every request carries the experiment secret,
compared in constant time; a wrong secret closes the connection without a reply, like the reference's server.

    {"type": "LOG", "secret": "..."}  ->  {"type": "OK", "ex_logs": str|None, "num_trials": N, "to_date": k,
                                            "stopped": s, "metric": best}
"""
from __future__ import annotations

import json
import secrets
import socket
import struct
import threading
from typing import Optional, Tuple


def _recv_exact(sock, n):
    buf = b""
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed the connection")
        buf += chunk
    return buf


def send_msg(sock, obj) -> None:
    raw = json.dumps(obj, default=str).encode()
    sock.sendall(struct.pack(">I", len(raw)) + raw)


def recv_msg(sock, max_bytes: int = 64 << 20):
    (n,) = struct.unpack(">I", _recv_exact(sock, 4))
    if n > max_bytes:
        raise ValueError(f"message of {n} bytes exceeds the {max_bytes}-byte limit")
    return json.loads(_recv_exact(sock, n).decode())


class LogServer:
    def __init__(self, driver, host: str = "127.0.0.1", port: int = 0, secret: Optional[str] = None):
        self.driver = driver
        self.secret = secret or secrets.token_hex(8)
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind((host, port))
        self.sock.listen(10)
        self.sock.settimeout(0.25)
        self.address: Tuple[str, int] = self.sock.getsockname()
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._serve, name="maggy-logserver", daemon=True)

    def start(self) -> "LogServer":
        self._thread.start()
        return self

    def _reply(self, msg: dict) -> dict:
        if msg.get("type") == "LOG":
            result, logs = self.driver.get_logs()
            result = result or {}
            return {"type": "OK", "ex_logs": logs or None, "num_trials": getattr(self.driver, "num_trials", 1),
                    "to_date": result.get("num_trials", 0), "stopped": result.get("early_stopped", 0),
                    "metric": result.get("best_val")}
        return {"type": "ERR"}   # unknown request type (reference: rpc.py:325-326)

    def _serve(self):
        while not self._stop.is_set():
            try:
                conn, _ = self.sock.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            with conn:
                conn.settimeout(5.0)
                try:
                    while True:
                        msg = recv_msg(conn, 1 << 16)
                        if not secrets.compare_digest(str(msg.get("secret", "")), self.secret):
                            break     # wrong secret: drop the connection, answer nothing
                        send_msg(conn, self._reply(msg))
                except (ConnectionError, socket.timeout, ValueError, OSError):
                    pass

    def stop(self):
        self._stop.set()
        try:
            self.sock.close()
        except OSError:
            pass
        if self._thread.is_alive():
            self._thread.join(timeout=1.0)


def poll_logs(address: Tuple[str, int], secret: str, timeout_s: float = 5.0) -> dict:
    """What the notebook side does: one LOG request, one reply."""
    with socket.create_connection(address, timeout=timeout_s) as s:
        send_msg(s, {"type": "LOG", "secret": secret})
        return recv_msg(s)
