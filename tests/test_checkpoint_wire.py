"""CPU test of the checkpoint wire format (Learner::Serialize / Parse).

The reference writes its checkpoints with protobuf-generated code: a stream of records, each a
`uint64` byte count followed by one proto2 message of mcmc/protos.proto (serialize.h:13-38).  The
host library encodes the same messages by hand.  Here every message kind is written by the host
library and compared BYTE FOR BYTE with what the protobuf runtime itself produces for the same
field values from the reference's schema, and bytes produced by the protobuf runtime are parsed
back by the host library -- so files interchange with the reference in both directions."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

import pymcmc

F = descriptor_pb2.FieldDescriptorProto
TYPES = {"bytes": F.TYPE_BYTES, "uint32": F.TYPE_UINT32, "uint64": F.TYPE_UINT64, "int32": F.TYPE_INT32,
         "double": F.TYPE_DOUBLE}
# the reference's mcmc/protos.proto:3-50 (all fields `required`, numbered in order)
SCHEMA = {
    "VectorStorage": [("bytes", "storage")],
    "RpmProperties": [("uint32", "rows"), ("uint32", "cols"), ("uint32", "rows_in_block")],
    "BetaProperties": [("uint32", "count_calls"), ("double", "theta_sum_time"), ("double", "grads_partial_time"),
                       ("double", "grads_sum_time"), ("double", "update_theta_time"), ("double", "normalize_time")],
    "PhiProperties": [("uint32", "count_calls"), ("double", "update_phi_time"), ("double", "update_pi_time")],
    "PerplexityProperties": [("uint32", "count_calls"), ("double", "ppx_time"), ("double", "accumulate_time")],
    "SampleStorage": [("bytes", "edges"), ("bytes", "nodes_vec"), ("uint32", "seed")],
    "LearnerProperties": [("uint32", "stepCount"), ("uint64", "time"), ("uint64", "samplingTime"), ("int32", "phase"),
                          ("double", "weight")],
}
KIND = {"BetaProperties": 0, "PhiProperties": 1, "PerplexityProperties": 2, "SampleStorage": 3,
        "LearnerProperties": 4, "VectorStorage": 5, "RpmProperties": 6}
REFERENCE_PROTO = "/root/reference/mcmc/protos.proto"


@pytest.fixture(scope="module")
def classes():
    fdp = descriptor_pb2.FileDescriptorProto(name="mcmc_protos.proto", package="mcmc", syntax="proto2")
    for name, fields in SCHEMA.items():
        m = fdp.message_type.add(name=name)
        for number, (ftype, fname) in enumerate(fields, 1):
            m.field.add(name=fname, number=number, type=TYPES[ftype], label=F.LABEL_REQUIRED)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return {name: message_factory.GetMessageClass(pool.FindMessageTypeByName("mcmc." + name)) for name in SCHEMA}


def test_schema_is_the_references_protos_proto():
    """the table above against the reference's own .proto text (present in the build container)"""
    if not os.path.exists(REFERENCE_PROTO):
        pytest.skip("/root/reference not present")
    text = open(REFERENCE_PROTO).read()
    for name, fields in SCHEMA.items():
        body = re.search(r"message\s+%s\s*\{(.*?)\}" % name, text, re.S).group(1)
        got = re.findall(r"required\s+(\w+)\s+(\w+)\s*=\s*(\d+)\s*;", body)
        assert [(t, n) for t, n, _ in got] == fields
        assert [int(k) for _, _, k in got] == list(range(1, len(fields) + 1))


def host_write(kind, ints=(), dbls=(), b1=b"", b2=b""):
    ia = np.array(list(ints) + [0] * 4, dtype=np.uint64)
    da = np.array(list(dbls) + [0.0] * 5, dtype=np.float64)
    out = C.create_string_buffer(len(b1) + len(b2) + 256)
    L = pymcmc.lib()
    L.mcmc_test_serialize.restype = C.c_uint64
    n = L.mcmc_test_serialize(kind, ia.ctypes.data_as(C.c_void_p), da.ctypes.data_as(C.c_void_p), b1,
                              C.c_uint64(len(b1)), b2, C.c_uint64(len(b2)), out, C.c_uint64(len(out)))
    assert n != 2 ** 64 - 1
    return out.raw[:n]


def host_parse(kind, record, cap1=0, cap2=0):
    ia, da = np.zeros(4, dtype=np.uint64), np.zeros(5, dtype=np.float64)
    b1, b2 = C.create_string_buffer(max(cap1, 1)), C.create_string_buffer(max(cap2, 1))
    n1, n2 = C.c_uint64(cap1), C.c_uint64(cap2)
    rc = pymcmc.lib().mcmc_test_parse(kind, record, C.c_uint64(len(record)), ia.ctypes.data_as(C.c_void_p),
                                      da.ctypes.data_as(C.c_void_p), b1, C.byref(n1), b2, C.byref(n2))
    return rc, ia, da, b1.raw[:n1.value], b2.raw[:n2.value]


def record(msg):
    payload = msg.SerializeToString()
    return struct.pack("<Q", len(payload)) + payload  # serialize.h:13-24


CASES = [
    ("BetaProperties", dict(count_calls=7, theta_sum_time=1.5, grads_partial_time=0.0, grads_sum_time=-2.25,
                            update_theta_time=1e300, normalize_time=3.0)),
    ("BetaProperties", dict(count_calls=0, theta_sum_time=0.0, grads_partial_time=0.0, grads_sum_time=0.0,
                            update_theta_time=0.0, normalize_time=0.0)),
    ("PhiProperties", dict(count_calls=4294967295, update_phi_time=12.125, update_pi_time=0.5)),
    ("PerplexityProperties", dict(count_calls=300, ppx_time=1e-9, accumulate_time=7.0)),
    ("SampleStorage", dict(edges=bytes(range(256)) * 9, nodes_vec=b"\x00\xff" * 70, seed=123456789)),
    ("SampleStorage", dict(edges=b"", nodes_vec=b"", seed=0)),
    ("LearnerProperties", dict(stepCount=1001, time=2 ** 63 + 5, samplingTime=129, phase=1, weight=27558.4)),
    ("LearnerProperties", dict(stepCount=1, time=0, samplingTime=2 ** 40, phase=-1, weight=-0.0)),
    ("VectorStorage", dict(storage=os.urandom(100000))),
    ("VectorStorage", dict(storage=b"")),
    ("RpmProperties", dict(rows=65608366, cols=512, rows_in_block=262144)),
]


@pytest.mark.parametrize("name,values", CASES)
def test_records_are_protobufs_bytes_and_parse_back(classes, name, values):
    msg = classes[name](**values)
    want = record(msg)
    ints = [v & (2 ** 64 - 1) for (t, f), v in zip(SCHEMA[name], [values[f] for _, f in SCHEMA[name]]) if t not in ("bytes", "double")]
    dbls = [values[f] for t, f in SCHEMA[name] if t == "double"]
    blobs = [values[f] for t, f in SCHEMA[name] if t == "bytes"] + [b"", b""]
    got = host_write(KIND[name], ints, dbls, blobs[0], blobs[1])
    assert got == want  # byte for byte what protoc-generated code writes
    # and the other direction: the protobuf runtime's bytes through the host parser
    rc, ia, da, b1, b2 = host_parse(KIND[name], want, len(blobs[0]), len(blobs[1]))
    assert rc == 0
    assert [int(x) for x in ia[:len(ints)]] == ints
    assert [float(x) for x in da[:len(dbls)]] == [float(x) for x in dbls]
    if name in ("SampleStorage", "VectorStorage"):
        assert b1 == blobs[0] and b2 == blobs[1]
    # the protobuf runtime reads the host library's record
    back = classes[name]()
    back.ParseFromString(got[8:])
    assert back == msg and struct.unpack("<Q", got[:8])[0] == len(got) - 8


def test_parse_rejects_a_buffer_of_the_wrong_size_and_truncation(classes):
    """serialize.h:62-69: a VectorStorage whose size differs from the destination buffer fails;
    so does a record cut short"""
    rec = record(classes["VectorStorage"](storage=b"x" * 64))
    assert host_parse(5, rec, 64)[0] == 0
    assert host_parse(5, rec, 63)[0] != 0 and host_parse(5, rec, 65)[0] != 0
    assert host_parse(5, rec[:-3], 64)[0] != 0
    lp = record(classes["LearnerProperties"](stepCount=1, time=2, samplingTime=3, phase=0, weight=1.0))
    assert host_parse(4, lp[:10])[0] != 0
