"""Build a protobuf FileDescriptorSet from the reference's four .proto schemas
(/root/reference/framework/model_parser/proto/{graph,node,tensor,operator}.proto) with a small .proto
parser (there is no protoc in this image) and store it as tests/golden/anakin_proto.desc.

tests/test_cpu_parser.py loads that descriptor set into google.protobuf and uses the resulting message
classes as the CANONICAL encoder / decoder the hand-written codecs (csrc/framework/graph.cpp,
anakin_b200/anakin_bin.py) are checked against, byte for byte. When /root/reference is present the test
re-derives the descriptors and checks the committed fixture is current.

  python tools/make_proto_descriptors.py            # writes the fixture
"""
import os
import re
import sys

from google.protobuf import descriptor_pb2 as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROTO_DIR = "/root/reference/framework/model_parser/proto"
FILES = ["operator.proto", "tensor.proto", "node.proto", "graph.proto"]
OUT = os.path.join(ROOT, "tests", "golden", "anakin_proto.desc")

SCALAR = {
    "double": D.FieldDescriptorProto.TYPE_DOUBLE, "float": D.FieldDescriptorProto.TYPE_FLOAT,
    "int32": D.FieldDescriptorProto.TYPE_INT32, "int64": D.FieldDescriptorProto.TYPE_INT64,
    "uint32": D.FieldDescriptorProto.TYPE_UINT32, "uint64": D.FieldDescriptorProto.TYPE_UINT64,
    "bool": D.FieldDescriptorProto.TYPE_BOOL, "string": D.FieldDescriptorProto.TYPE_STRING,
    "bytes": D.FieldDescriptorProto.TYPE_BYTES,
}


def _tokens(text):
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"[A-Za-z_][\w.]*|\d+|\"[^\"]*\"|[{}=;<>,\[\]]", text)


class _Parser:
    def __init__(self, name, text, known_enums, known_msgs):
        self.t = _tokens(text)
        self.i = 0
        self.fd = D.FileDescriptorProto(name=name, syntax="proto3")
        self.enums, self.msgs = known_enums, known_msgs

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self, want=None):
        tok = self.t[self.i]
        self.i += 1
        if want is not None and tok != want:
            raise SyntaxError("expected %r, got %r" % (want, tok))
        return tok

    def parse(self):
        while self.peek() is not None:
            tok = self.take()
            if tok == "syntax":
                self.take("="); self.take(); self.take(";")
            elif tok == "import":
                self.fd.dependency.append(self.take().strip('"')); self.take(";")
            elif tok == "message":
                self.message(self.fd.message_type.add(), "")
            elif tok == "enum":
                self.enum(self.fd.enum_type.add(), "")
            elif tok == ";":
                pass
            else:
                raise SyntaxError("unexpected top-level token %r" % tok)
        return self.fd

    def enum(self, e, scope):
        e.name = self.take()
        self.enums.add(scope + e.name)
        self.take("{")
        while self.peek() != "}":
            name = self.take()
            self.take("=")
            e.value.add(name=name, number=int(self.take()))
            self.take(";")
        self.take("}")

    def field(self, m, scope, label_repeated, ftype, oneof_index=None):
        name = self.take()
        self.take("=")
        number = int(self.take())
        self.take(";")
        f = m.field.add(name=name, number=number, json_name=name)
        f.label = D.FieldDescriptorProto.LABEL_REPEATED if label_repeated else D.FieldDescriptorProto.LABEL_OPTIONAL
        if ftype in SCALAR:
            f.type = SCALAR[ftype]
        else:
            f.type_name = ftype      # resolved after every file is parsed
        if oneof_index is not None:
            f.oneof_index = oneof_index
        return f

    def message(self, m, scope):
        m.name = self.take()
        self.msgs.add(scope + m.name)
        inner = scope + m.name + "."
        self.take("{")
        while self.peek() != "}":
            tok = self.take()
            if tok == "message":
                self.message(m.nested_type.add(), inner)
            elif tok == "enum":
                self.enum(m.enum_type.add(), inner)
            elif tok == "oneof":
                idx = len(m.oneof_decl)
                m.oneof_decl.add(name=self.take())
                self.take("{")
                while self.peek() != "}":
                    self.field(m, inner, False, self.take(), oneof_index=idx)
                self.take("}")
            elif tok == "map":
                self.take("<"); kt = self.take(); self.take(","); vt = self.take(); self.take(">")
                name = self.take(); self.take("="); number = int(self.take()); self.take(";")
                entry = m.nested_type.add(name="".join(p.capitalize() for p in name.split("_")) + "Entry")
                entry.options.map_entry = True
                k = entry.field.add(name="key", number=1, json_name="key", label=D.FieldDescriptorProto.LABEL_OPTIONAL)
                k.type = SCALAR[kt]
                v = entry.field.add(name="value", number=2, json_name="value", label=D.FieldDescriptorProto.LABEL_OPTIONAL)
                if vt in SCALAR:
                    v.type = SCALAR[vt]
                else:
                    v.type_name = vt
                f = m.field.add(name=name, number=number, json_name=name, label=D.FieldDescriptorProto.LABEL_REPEATED,
                                type=D.FieldDescriptorProto.TYPE_MESSAGE)
                f.type_name = "." + inner + entry.name
                self.msgs.add(inner + entry.name)
            elif tok == "repeated":
                self.field(m, inner, True, self.take())
            elif tok == ";":
                pass
            else:
                self.field(m, inner, False, tok)
        self.take("}")
        if self.peek() == ";":
            self.take()


def _resolve(fds, enums, msgs):
    def fix(m, scope):
        for f in m.field:
            if f.type_name and not f.type_name.startswith("."):
                # innermost scope first, then file scope (the four files use no packages)
                cands = [scope + m.name + "." + f.type_name, scope + f.type_name, f.type_name]
                hit = next((c for c in cands if c in enums or c in msgs), None)
                if hit is None:
                    raise NameError("unresolved type %s in %s" % (f.type_name, scope))
                f.type = D.FieldDescriptorProto.TYPE_ENUM if hit in enums else D.FieldDescriptorProto.TYPE_MESSAGE
                f.type_name = "." + hit
        for n in m.nested_type:
            fix(n, scope + m.name + ".")
    for fd in fds:
        for m in fd.message_type:
            fix(m, "")


def build_descriptor_set(proto_dir=PROTO_DIR):
    enums, msgs, fds = set(), set(), []
    for fn in FILES:
        with open(os.path.join(proto_dir, fn)) as f:
            fds.append(_Parser(fn, f.read(), enums, msgs).parse())
    _resolve(fds, enums, msgs)
    s = D.FileDescriptorSet()
    s.file.extend(fds)
    return s


def message_classes(desc_bytes):
    """{'GraphProto': class, ...} from a serialized FileDescriptorSet."""
    from google.protobuf import descriptor_pool, message_factory
    s = D.FileDescriptorSet()
    s.ParseFromString(desc_bytes)
    pool = descriptor_pool.DescriptorPool()
    for fd in s.file:
        pool.Add(fd)
    out = {}
    for fd in s.file:
        for m in fd.message_type:
            out[m.name] = message_factory.GetMessageClass(pool.FindMessageTypeByName(m.name))
    return out


if __name__ == "__main__":
    ds = build_descriptor_set()
    blob = ds.SerializeToString(deterministic=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(blob)
    cls = message_classes(blob)
    print("wrote %s (%d bytes); messages: %s" % (OUT, len(blob), ", ".join(sorted(cls))))
