"""The reference's DataFrame / logical-plan API and its query planner, over the operator vocabulary of an engine.

north_star: "Expression.evaluate, PhysicalPlan.execute and the DataFrame/logical-plan API all stay unchanged". On a JVM
that is the Kotlin shim of INTEGRATION.md; here (no JDK) this module restates that layer of kquerydiy/src/Main.kt in the
host language the tests use, with the reference's names and error behaviour, so that its queries run unchanged:

    Schema / Field (Main.kt:29-54)             LogicalPlan: Scan, Projection, Aggregate (68-166) [+ Selection]
    LogicalExpr: Column, ColumnIndex, Alias, CastExpr, Max (72-101, 421-440, 1207-1216) [+ Sum, Min, Count, literals, binary]
    DataFrame.project/aggregate/schema/logicalPlan (359-385) [+ filter]
    ExecutionContext.sql/registerCsv/registerDataSource/execute (387-419)
    ProjectionPushDownRule (707-770), createPhysicalExpr / createPhysicalPlan (662-706)
    ScanExec, ProjectionExec, HashAggregateExec (564-660) [+ SelectionExec, fused into its parent where one kernel does both]
    DataSource: CsvDataSource (276-357), InMemoryDataSource (1292-1304)

`engine` is the object that runs the operators: kqgpu.Engine(ctx). This module only needs its vocabulary — col, cast,
lit_*, binary, project, filter, filter_project, HashAggregate, csv_header, csv_scan / csv_batches — and imports nothing
else, which is what lets the test-suite hand it a checker with the same vocabulary and verify this layer without a
device; there is no fallback in here: nothing in this module computes on data (planning, schemas and names only), and
without an engine nothing runs. What the reference lacks is marked [+ ...] above.
"""
from __future__ import annotations

F64, UTF8, I64, BOOL, DATE32, I32 = 1, 2, 3, 4, 5, 6


# ---- the exception classes the reference throws (kq_status keeps the same names, include/kqgpu.h) -----------------------
class SQLException(Exception):                     # Main.kt:79, 666, 1218-1290
    pass


class IllegalStateException(Exception):            # Main.kt:677, 696, 704, 1108 ...
    pass


class IllegalArgumentException(Exception):         # Schema.select, Main.kt:49
    pass


class UnsupportedOperationException(Exception):    # Main.kt:767
    pass


# ---- schema -------------------------------------------------------------------------------------------------------------
class ArrowType:
    """ArrowTypes.DoubleType / StringType (Main.kt:19-22) and the extension types of rule E1; prints like Arrow Java."""

    _NAMES = {F64: "FloatingPoint(DOUBLE)", UTF8: "Utf8", I64: "Int(64, true)", BOOL: "Bool", DATE32: "Date(DAY)", I32: "Int(32, true)"}

    def __init__(self, kq_type: int):
        if kq_type not in self._NAMES:
            raise IllegalStateException(f"Unsupported data type: {kq_type}")          # Main.kt:195
        self.kq_type = kq_type

    def __eq__(self, other):
        return isinstance(other, ArrowType) and other.kq_type == self.kq_type

    def __hash__(self):
        return hash(self.kq_type)

    def __repr__(self):
        return self._NAMES[self.kq_type]


DoubleType, StringType = ArrowType(F64), ArrowType(UTF8)
Int64Type, BooleanType, Date32Type = ArrowType(I64), ArrowType(BOOL), ArrowType(DATE32)


class Field:
    def __init__(self, name: str, dataType: ArrowType):
        self.name, self.dataType = name, dataType

    def __eq__(self, other):
        return isinstance(other, Field) and (self.name, self.dataType) == (other.name, other.dataType)

    def __repr__(self):
        return f"Field(name={self.name}, dataType={self.dataType})"


class Schema:
    def __init__(self, fields):
        self.fields = list(fields)

    def select(self, names):
        """Main.kt:41-53: every name must match exactly one field."""
        out = []
        for name in names:
            m = [f for f in self.fields if f.name == name]
            if len(m) != 1:
                raise IllegalArgumentException(f"Field {name} not found")
            out.append(m[0])
        return Schema(out)

    def names(self):
        return [f.name for f in self.fields]

    def __eq__(self, other):
        return isinstance(other, Schema) and self.fields == other.fields

    def __repr__(self):
        return f"Schema(fields={self.fields})"


# ---- logical expressions ------------------------------------------------------------------------------------------------
class LogicalExpr:
    def toField(self, input: "LogicalPlan") -> Field:
        raise NotImplementedError


class Column(LogicalExpr):
    def __init__(self, name: str):
        self.name = name

    def toField(self, input):
        for f in input.schema().fields:
            if f.name == self.name:
                return f
        raise SQLException(f"No column named '{self.name}'")

    def __repr__(self):
        return f"#{self.name}"


class ColumnIndex(LogicalExpr):
    def __init__(self, i: int):
        self.i = i

    def toField(self, input):
        return input.schema().fields[self.i]

    def __repr__(self):
        return f"#{self.i}"


class Alias(LogicalExpr):
    def __init__(self, expr: LogicalExpr, alias: str):
        self.expr, self.alias = expr, alias

    def toField(self, input):
        return Field(self.alias, self.expr.toField(input).dataType)

    def __repr__(self):
        return f"{self.expr} as {self.alias}"


class CastExpr(LogicalExpr):
    def __init__(self, expr: LogicalExpr, dataType: ArrowType):
        self.expr, self.dataType = expr, dataType

    def toField(self, input):
        return Field(self.expr.toField(input).name, self.dataType)

    def __repr__(self):
        return f"CAST({self.expr} AS {self.dataType})"


class AggregateExpr(LogicalExpr):
    def __init__(self, name: str, expr: LogicalExpr):
        self.name, self.expr = name, expr

    def toField(self, input):
        return Field(self.name, self.expr.toField(input).dataType)

    def __repr__(self):
        return f"{self.name}({self.expr})"


class Max(AggregateExpr):
    def __init__(self, input: LogicalExpr):
        super().__init__("MAX", input)


class Min(AggregateExpr):                          # [+] rule E5
    def __init__(self, input: LogicalExpr):
        super().__init__("MIN", input)


class Sum(AggregateExpr):                          # [+] rule E6
    def __init__(self, input: LogicalExpr):
        super().__init__("SUM", input)


class Count(AggregateExpr):                        # [+] rule E7: non-null rows, Int64
    def __init__(self, input: LogicalExpr):
        super().__init__("COUNT", input)

    def toField(self, input):
        return Field("COUNT", Int64Type)


class Literal(LogicalExpr):                        # [+] rule E1
    def __init__(self, value, dataType: ArrowType):
        self.value, self.dataType = value, dataType

    def toField(self, input):
        return Field(str(self.value), self.dataType)

    def __repr__(self):
        return f"'{self.value}'" if self.dataType == StringType else str(self.value)


def lit(value) -> Literal:
    if isinstance(value, bool):
        return Literal(value, BooleanType)
    if isinstance(value, int):
        return Literal(value, Int64Type)
    if isinstance(value, float):
        return Literal(value, DoubleType)
    if isinstance(value, str):
        return Literal(value, StringType)
    import datetime
    if isinstance(value, datetime.date):
        return Literal((value - datetime.date(1970, 1, 1)).days, Date32Type)
    raise IllegalStateException(f"Unsupported literal: {value!r}")


class BinaryExpr(LogicalExpr):                     # [+] rules E2-E4; names and symbols as in KQuery
    OPS = {"EQ": ("eq", "="), "NE": ("neq", "!="), "LT": ("lt", "<"), "LE": ("lteq", "<="), "GT": ("gt", ">"), "GE": ("gteq", ">="),
           "AND": ("and", "AND"), "OR": ("or", "OR"), "ADD": ("add", "+"), "SUB": ("subtract", "-"), "MUL": ("mult", "*"), "DIV": ("div", "/")}

    def __init__(self, op: str, l: LogicalExpr, r: LogicalExpr):
        if op not in self.OPS:
            raise IllegalStateException(f"Unknown binary operator: {op}")
        self.op, self.l, self.r = op, l, r

    def toField(self, input):
        name = self.OPS[self.op][0]
        if self.op in ("ADD", "SUB", "MUL", "DIV"):
            return Field(name, self.l.toField(input).dataType)
        return Field(name, BooleanType)

    def __repr__(self):
        return f"{self.l} {self.OPS[self.op][1]} {self.r}"


def col(name: str) -> Column:
    return Column(name)


def cast(expr: LogicalExpr, dataType: ArrowType) -> CastExpr:
    return CastExpr(expr, dataType)


# ---- logical plans ------------------------------------------------------------------------------------------------------
class LogicalPlan:
    def schema(self) -> Schema:
        raise NotImplementedError

    def children(self):
        raise NotImplementedError


class Scan(LogicalPlan):
    def __init__(self, path: str, dataSource: "DataSource", projection):
        self.path, self.dataSource, self.projection = path, dataSource, list(projection)
        s = dataSource.schema()
        self._schema = s if not self.projection else s.select(self.projection)

    def schema(self):
        return self._schema

    def children(self):
        return []

    def __repr__(self):
        return f"Scan: {self.path}; projection=" + ("None" if not self.projection else "[" + ", ".join(self.projection) + "]")


class Projection(LogicalPlan):
    def __init__(self, input: LogicalPlan, expr):
        self.input, self.expr = input, list(expr)

    def schema(self):
        return Schema([e.toField(self.input) for e in self.expr])

    def children(self):
        return [self.input]

    def __repr__(self):
        return "Projection: " + ",".join(str(e) for e in self.expr)


class Selection(LogicalPlan):                      # [+] the Selection node the reference's plan lacks
    def __init__(self, input: LogicalPlan, expr: LogicalExpr):
        self.input, self.expr = input, expr

    def schema(self):
        return self.input.schema()

    def children(self):
        return [self.input]

    def __repr__(self):
        return f"Selection: {self.expr}"


class Aggregate(LogicalPlan):
    def __init__(self, input: LogicalPlan, groupExpr, aggExpr):
        self.input, self.groupExpr, self.aggExpr = input, list(groupExpr), list(aggExpr)

    def schema(self):
        return Schema([e.toField(self.input) for e in self.groupExpr] + [e.toField(self.input) for e in self.aggExpr])

    def children(self):
        return [self.input]

    def __repr__(self):
        return f"Aggregate: groupExpr={self.groupExpr}, aggregateExpr={self.aggExpr}"


def format_plan(plan, indent: int = 0) -> str:
    """The plan as an indented tree, one node per line (logical or physical)."""
    out = "\t" * indent + str(plan) + "\n"
    for c in plan.children():
        out += format_plan(c, indent + 1)
    return out


# ---- DataFrame and ExecutionContext ---------------------------------------------------------------------------------------
class DataFrame:
    """DataFrameImpl (Main.kt:366-385)."""

    def __init__(self, plan: LogicalPlan):
        self._plan = plan

    def project(self, expr) -> "DataFrame":
        return DataFrame(Projection(self._plan, expr))

    def filter(self, expr: LogicalExpr) -> "DataFrame":        # [+]
        return DataFrame(Selection(self._plan, expr))

    def aggregate(self, groupBy, aggregateExpr) -> "DataFrame":
        return DataFrame(Aggregate(self._plan, groupBy, aggregateExpr))

    def schema(self) -> Schema:
        return self._plan.schema()

    def logicalPlan(self) -> LogicalPlan:
        return self._plan


class ExecutionContext:
    """ExecutionContext (Main.kt:387-419), bound to the engine that runs the physical operators."""

    def __init__(self, engine, merge=None):
        """merge: None, "allreduce" or "repartition" — when every rank of a communicator (kq_comm_init on the engine's
        context) runs the same plan over its shard of the rows, the aggregate merges the partial tables with that collective
        after draining its input: main()'s partition -> partial -> merge (Main.kt:1309-1325) inside the operator instead of a
        second query. Every rank must execute the plan (the merges are collectives)."""
        if merge not in (None, "allreduce", "repartition"):
            raise IllegalArgumentException(f"unknown merge {merge!r}")
        self.engine, self.merge = engine, merge
        self._tables = {}

    def sql(self, sql: str) -> DataFrame:
        from . import sql as _sql
        ast = _sql.SqlParser(_sql.SqlTokenizer(sql).tokenize()).parse()
        if not isinstance(ast, _sql.SqlSelect):
            raise IllegalStateException(f"Expected a SELECT statement, found {ast}")
        return DataFrame(_sql.createDataFrame(ast, self._tables).logicalPlan())

    def csv(self, filename: str, hasHeaders: bool = True, batchSize: int = 1000) -> DataFrame:
        return DataFrame(Scan(filename, CsvDataSource(self.engine, filename, hasHeaders, batchSize), []))

    def register(self, tablename: str, df: DataFrame):
        self._tables[tablename] = df

    def registerCsv(self, tablename: str, filename: str):
        self.register(tablename, self.csv(filename))

    def registerDataSource(self, tablename: str, datasource: "DataSource"):
        self.register(tablename, DataFrame(Scan(tablename, datasource, [])))

    def execute(self, df):
        """Sequence<RecordBatch> of the engine's batches (Main.kt:411-419): optimise, plan, run."""
        plan = df.logicalPlan() if isinstance(df, DataFrame) else df
        return createPhysicalPlan(ProjectionPushDownRule().optimize(plan), self.engine, self.merge).execute()


# ---- data sources ---------------------------------------------------------------------------------------------------------
class DataSource:
    def schema(self) -> Schema:
        raise NotImplementedError

    def scan(self, projection):
        raise NotImplementedError


class CsvDataSource(DataSource):
    """CsvDataSource (Main.kt:276-357) with the tokenising done by the engine: every column is Utf8 (345-349). `source` is a
    file name, or the file's bytes. With an engine that has csv_batches (the GPU) the text streams through the reader and
    `batchSize` rows is a hint the device ignores — its batches are pieces of text (rule R11: results do not depend on
    batch boundaries); otherwise the whole text is one batch."""

    def __init__(self, engine, source, hasHeaders: bool = True, batchSize: int = 1000, piece_bytes: int = 0):
        self.engine, self.source, self.hasHeaders, self.batchSize, self.piece_bytes = engine, source, hasHeaders, batchSize, piece_bytes
        self._text = None
        self._schema = None

    def _bytes(self) -> bytes:
        if self._text is None:
            if isinstance(self.source, (bytes, bytearray)):
                self._text = bytes(self.source)
            else:
                import os
                if not os.path.exists(self.source):
                    raise FileNotFoundError(os.path.abspath(self.source))                 # Main.kt:306-309
                with open(self.source, "rb") as f:
                    self._text = f.read()
        return self._text

    def schema(self):
        if self._schema is None:
            names, _ = self.engine.csv_header(self._bytes(), self.hasHeaders)
            self._schema = Schema([Field(n, StringType) for n in names])
        return self._schema

    def scan(self, projection):
        projection = list(projection)
        self.schema().select(projection)                     # unknown names throw here, like Main.kt:311-315
        text = self._bytes()
        if hasattr(self.engine, "csv_batches"):
            yield from self.engine.csv_batches(text, self.hasHeaders, projection or None, piece_bytes=self.piece_bytes)
        else:
            b = self.engine.csv_scan(text, self.hasHeaders, projection or None)
            if b.row_count() > 0:                            # Main.kt:245-247: no batch without rows
                yield b


class InMemoryDataSource(DataSource):
    """InMemoryDataSource (Main.kt:1292-1304): batches already held by the engine. scan() re-indexes the columns by the
    projected names; an empty projection means all columns."""

    def __init__(self, engine, schema: Schema, data):
        self.engine, self._schema, self.data = engine, schema, list(data)
        for batch in self.data:                     # a schema that lies about its columns fails here, not inside a kernel
            if batch.num_columns() != len(schema.fields):
                raise IllegalStateException(f"batch with {batch.num_columns()} columns under a schema of {len(schema.fields)} fields")
            for i, f in enumerate(schema.fields):
                if batch.field(i).type() != f.dataType.kq_type:
                    raise IllegalStateException(f"column {f.name}: the batch holds {ArrowType(batch.field(i).type())}, the schema says {f.dataType}")

    def schema(self):
        return self._schema

    def scan(self, projection):
        projection = list(projection)
        if not projection:
            yield from self.data
            return
        names = self._schema.names()
        idx = []
        for name in projection:
            if name not in names:
                raise IllegalArgumentException(f"Field {name} not found")
            idx.append(names.index(name))
        for batch in self.data:
            yield self.engine.RecordBatch.from_columns([batch.field(i) for i in idx], batch.row_count())


# ---- optimizer --------------------------------------------------------------------------------------------------------------
def extractColumns(expr, input: LogicalPlan, accum: set):
    """Main.kt:712-737. The reference's branch for a nested AggregateExpr adds the literal name "fare_amount" (a leftover of
    its taxi-data main()); here the aggregate's own input expression is followed instead."""
    if isinstance(expr, (list, tuple)):
        for e in expr:
            extractColumns(e, input, accum)
    elif isinstance(expr, Column):
        accum.add(expr.name)
    elif isinstance(expr, ColumnIndex):
        accum.add(input.schema().fields[expr.i].name)
    elif isinstance(expr, (Alias, CastExpr, AggregateExpr)):
        extractColumns(expr.expr, input, accum)
    elif isinstance(expr, BinaryExpr):
        extractColumns(expr.l, input, accum)
        extractColumns(expr.r, input, accum)
    elif isinstance(expr, Literal):
        pass
    else:
        raise IllegalStateException(f"extractColumns does not support expression: {expr}")


class ProjectionPushDownRule:
    """Main.kt:739-770: collect the column names a plan refers to on the way down and hand them to the Scan, sorted."""

    def optimize(self, plan: LogicalPlan) -> LogicalPlan:
        return self._pushDown(plan, set())

    def _pushDown(self, plan, columnNames: set):
        if isinstance(plan, Projection):
            extractColumns(plan.expr, plan.input, columnNames)
            return Projection(self._pushDown(plan.input, columnNames), plan.expr)
        if isinstance(plan, Selection):
            extractColumns(plan.expr, plan.input, columnNames)
            return Selection(self._pushDown(plan.input, columnNames), plan.expr)
        if isinstance(plan, Aggregate):
            extractColumns(plan.groupExpr, plan.input, columnNames)
            extractColumns([a.expr for a in plan.aggExpr], plan.input, columnNames)
            return Aggregate(self._pushDown(plan.input, columnNames), plan.groupExpr, plan.aggExpr)
        if isinstance(plan, Scan):
            valid = plan.dataSource.schema().names()
            return Scan(plan.path, plan.dataSource, sorted(n for n in set(valid) if n in columnNames))
        raise UnsupportedOperationException(str(plan))


# ---- physical planning --------------------------------------------------------------------------------------------------------
_LITERALS = {F64: "lit_f64", UTF8: "lit_utf8", I64: "lit_i64", BOOL: "lit_bool", DATE32: "lit_date32"}


def createPhysicalExpr(expr: LogicalExpr, input: LogicalPlan, engine):
    """Main.kt:662-678: logical expression -> the engine's expression over the columns of `input`."""
    if isinstance(expr, Column):
        names = input.schema().names()
        if expr.name not in names:
            raise SQLException(f"No column named '{expr.name}'")
        return engine.col(names.index(expr.name))
    if isinstance(expr, ColumnIndex):
        return engine.col(expr.i)
    if isinstance(expr, Alias):
        return createPhysicalExpr(expr.expr, input, engine)
    if isinstance(expr, CastExpr):
        return engine.cast(createPhysicalExpr(expr.expr, input, engine), expr.dataType.kq_type)
    if isinstance(expr, Literal):
        return getattr(engine, _LITERALS[expr.dataType.kq_type])(expr.value)
    if isinstance(expr, BinaryExpr):
        return engine.binary(expr.op, createPhysicalExpr(expr.l, input, engine), createPhysicalExpr(expr.r, input, engine))
    raise IllegalStateException(f"Unknown expr: {expr}")


class PhysicalPlan:
    def schema(self) -> Schema:
        raise NotImplementedError

    def execute(self):
        raise NotImplementedError

    def children(self):
        raise NotImplementedError


class ScanExec(PhysicalPlan):
    def __init__(self, ds: DataSource, projection):
        self.ds, self.projection = ds, list(projection)

    def schema(self):
        return self.ds.schema().select(self.projection) if self.projection else self.ds.schema()

    def execute(self):
        return self.ds.scan(self.projection)

    def children(self):
        return []

    def __repr__(self):
        return f"ScanExec: schema={self.schema()}, projection={self.projection}"


class ProjectionExec(PhysicalPlan):
    """ProjectionExec (Main.kt:582-603): one output batch per input batch. With a predicate it is the fused
    filter+projection kernel (one pass, ordered compaction) — a Selection directly below a Projection."""

    def __init__(self, engine, input: PhysicalPlan, schema: Schema, expr, predicate=None, bare=None):
        self.engine, self.input, self._schema, self.expr, self.predicate = engine, input, schema, list(expr), predicate
        self.bare = list(bare) if bare is not None else [False] * len(self.expr)      # expr k is a column reference: an alias, or a gather below a filter

    def schema(self):
        return self._schema

    def execute(self):
        for batch in self.input.execute():
            if self.predicate is None:
                yield self.engine.project(self.expr, batch)
            else:
                yield self.engine.filter_project(self.predicate, self.expr, batch)

    def children(self):
        return [self.input]

    def __repr__(self):
        return f"ProjectionExec: {len(self.expr)} expressions" + (" (fused with the selection below)" if self.predicate is not None else "")


class SelectionExec(PhysicalPlan):                 # [+] FilterExec: surviving rows of every column, in input order
    def __init__(self, engine, input: PhysicalPlan, predicate):
        self.engine, self.input, self.predicate = engine, input, predicate

    def schema(self):
        return self.input.schema()

    def execute(self):
        for batch in self.input.execute():
            yield self.engine.filter(self.predicate, batch)

    def children(self):
        return [self.input]

    def __repr__(self):
        return "SelectionExec"


class HashAggregateExec(PhysicalPlan):
    """HashAggregateExec (Main.kt:605-660): drains its input, then yields exactly ONE batch — group columns, then aggregates
    (rule R10; zero input rows give one batch without rows). With a predicate the filter runs inside the aggregate kernel."""

    def __init__(self, engine, input: PhysicalPlan, groupExpr, aggregateExpr, schema: Schema, predicate=None, merge=None):
        self.engine, self.input, self.groupExpr, self.aggregateExpr, self._schema, self.predicate = engine, input, list(groupExpr), list(aggregateExpr), schema, predicate
        self.merge = merge          # "allreduce" / "repartition": the collective merge of the ranks' partial tables (ExecutionContext)

    def schema(self):
        return self._schema

    def execute(self):
        agg = self.engine.HashAggregate(self.groupExpr, self.aggregateExpr, pred=self.predicate)
        fed = False
        for batch in self.input.execute():
            agg.update(batch)
            fed = True
        if fed:
            if self.merge == "allreduce":
                agg.merge_allreduce()               # afterwards every rank holds the complete result
            elif self.merge == "repartition":
                agg.repartition_alltoall()          # afterwards every key lives on exactly one rank
            yield agg.finalize()
        else:       # no input batch at all: the engine never saw a column type; the plan knows them
            import pyarrow as pa
            types = {F64: pa.float64(), UTF8: pa.string(), I64: pa.int64(), BOOL: pa.bool_(), DATE32: pa.date32(), I32: pa.int32()}
            yield self.engine.RecordBatch.from_arrow([pa.array([], type=types[f.dataType.kq_type]) for f in self._schema.fields], 0)

    def children(self):
        return [self.input]

    def __repr__(self):
        return (f"HashAggregateExec: {len(self.groupExpr)} group expressions, aggregates {[k for k, _ in self.aggregateExpr]}"
                + (" (fused with the selection below)" if self.predicate is not None else ""))


def createPhysicalPlan(plan: LogicalPlan, engine, merge=None) -> PhysicalPlan:
    """Main.kt:680-706 — the one place where operators are chosen. A Selection directly below a Projection or an
    Aggregate is folded into that operator's kernel instead of materialising the filtered batch."""
    if isinstance(plan, Scan):
        return ScanExec(plan.dataSource, plan.projection)
    if isinstance(plan, Selection):
        return SelectionExec(engine, createPhysicalPlan(plan.input, engine, merge), createPhysicalExpr(plan.expr, plan.input, engine))
    if isinstance(plan, (Projection, Aggregate)):
        source, predicate = plan.input, None
        if isinstance(source, Selection):
            predicate = createPhysicalExpr(source.expr, source.input, engine)
            source = source.input
        input = createPhysicalPlan(source, engine, merge)
        if isinstance(plan, Projection):
            expr = [createPhysicalExpr(e, plan.input, engine) for e in plan.expr]
            def bare(e):
                return bare(e.expr) if isinstance(e, Alias) else isinstance(e, (Column, ColumnIndex))
            return ProjectionExec(engine, input, Schema([e.toField(plan.input) for e in plan.expr]), expr, predicate, [bare(e) for e in plan.expr])
        groupExpr = [createPhysicalExpr(e, plan.input, engine) for e in plan.groupExpr]
        aggregateExpr = []
        for a in plan.aggExpr:
            if not isinstance(a, (Max, Min, Sum, Count)):
                raise IllegalStateException(f"Unsupported aggregate function: {a}")      # Main.kt:696
            aggregateExpr.append((a.name, createPhysicalExpr(a.expr, plan.input, engine)))
        return HashAggregateExec(engine, input, groupExpr, aggregateExpr, plan.schema(), predicate, merge)
    raise IllegalStateException("Unknown physical plan")


def explain(plan: PhysicalPlan, nullable: bool = False, compile: bool = True):
    """The CUDA source the engine generates for every operator of a physical plan — [(operator, source), ...], children
    first — compiled for sm_100a when `compile` (NVRTC needs no device: with kqgpu.Exprs() as the engine this runs on a CPU
    box, from the query string to the cubin). `nullable`: whether the scanned columns may hold nulls (CSV columns never do,
    rule C6). Operators without a kernel of their own (ScanExec) are left out."""
    out = []
    for c in plan.children():
        out += explain(c, nullable, compile)
    if isinstance(plan, (ProjectionExec, SelectionExec, HashAggregateExec)):
        fields = plan.input.schema().fields
        types, nulls = [f.dataType.kq_type for f in fields], [int(nullable)] * len(fields)
        # a bare column reference is not kernel work: without a filter it is an alias of the input column (rule R4), below a
        # filter a fixed-width column is compacted by the kernel and a Utf8 column is gathered by its selection vector
        if isinstance(plan, ProjectionExec):
            exprs = [e for e, b, f in zip(plan.expr, plan.bare, plan.schema().fields)
                     if not b or (plan.predicate is not None and f.dataType.kq_type != UTF8)]
            if not exprs and plan.predicate is None:
                return out                                   # only aliases: no kernel at all
            src = plan.engine.explain_filter_project(plan.predicate, exprs, types, nulls, compile)
        elif isinstance(plan, SelectionExec):
            src = plan.engine.explain_filter_project(plan.predicate, [plan.engine.col(i) for i, t in enumerate(types) if t != UTF8], types, nulls, compile)
        else:
            src = plan.engine.explain_hashagg(plan.groupExpr, plan.aggregateExpr, types, nulls, plan.predicate, compile)
        out.append((plan, src))
    return out


def printQueryResult(queryResult, file=None):
    """Main.kt:1344-1353: every row of every batch, cells separated by a blank (the batches are downloaded here)."""
    for batch in queryResult:
        cols = [a.to_pylist() for a in batch.to_arrow()]
        for row in zip(*cols):
            print(" ".join(str(v) for v in row) + " ", file=file)
