"""The reference's SQL front-end (kquerydiy/src/Main.kt:807-1290) — tokenizer, Pratt parser, planner — restated, plus the
extensions SURVEY.md §8 f4 names so that BASELINE.json's query strings run through ExecutionContext.sql unchanged:

    reference : SELECT exprs FROM table [GROUP BY exprs] [ORDER BY exprs];  identifiers, `quoted identifiers`, CAST(x AS double),
                MAX(x), x AS alias.  ORDER BY is parsed and ignored (there is no Sort operator, Main.kt:1276-1279).
    [+]       : WHERE; SUM / MIN / COUNT; long, double and string literals; = != <> < <= > >= AND OR + - * / with the usual
                precedences and parentheses; DATE 'yyyy-mm-dd'; a select list without aggregates plans a plain projection.

Errors follow the reference: SQLException for unknown tables, columns, data types and functions, IllegalStateException for
tokens the parser does not expect (Main.kt:1108, 1129, 1136, 1181, 1218-1290).
"""
from __future__ import annotations

from .plan import (AggregateExpr, Alias, BinaryExpr, CastExpr, Column, ColumnIndex, Count, DataFrame, DoubleType,
                   IllegalStateException, Int64Type, LogicalExpr, Max, Min, SQLException, StringType, Sum, lit)

KEYWORDS = {"AS", "BY", "CAST", "DOUBLE", "FROM", "GROUP", "MAX", "ORDER", "SELECT",          # Main.kt:807-822
            "WHERE", "AND", "OR", "SUM", "MIN", "COUNT", "ASC", "DESC", "DATE"}                   # [+]
SYMBOLS = ["<=", ">=", "!=", "<>", "(", ")", ",", "=", "<", ">", "+", "-", "*", "/"]               # longest first; ( ) , are the reference's


class TokenizeException(Exception):                # Main.kt:1046
    pass


class Token:
    def __init__(self, text: str, type: str, endOffset: int):
        self.text, self.type, self.endOffset = text, type, endOffset          # type: KEYWORD, SYMBOL, LONG, DOUBLE, STRING, IDENTIFIER

    def __repr__(self):
        return f'Token("{self.text}", {self.type}, {self.endOffset})'


class TokenStream:
    def __init__(self, tokens):
        self.tokens, self.i = list(tokens), 0

    def peek(self):
        return self.tokens[self.i] if self.i < len(self.tokens) else None

    def next(self):
        t = self.peek()
        if t is not None:
            self.i += 1
        return t

    def consumeKeyword(self, s: str) -> bool:
        t = self.peek()
        if t is not None and t.type == "KEYWORD" and t.text.upper() == s:
            self.i += 1
            return True
        return False

    def consumeKeywords(self, words) -> bool:
        save = self.i
        for w in words:
            if not self.consumeKeyword(w):
                self.i = save
                return False
        return True

    def consumeSymbol(self, s: str) -> bool:
        t = self.peek()
        if t is not None and t.type == "SYMBOL" and t.text == s:
            self.i += 1
            return True
        return False


class SqlTokenizer:
    def __init__(self, sql: str):
        self.sql, self.offset = sql, 0

    def tokenize(self) -> TokenStream:
        out = []
        while True:
            t = self._next()
            if t is None:
                return TokenStream(out)
            out.append(t)

    def _next(self):
        s = self.sql
        while self.offset < len(s) and s[self.offset].isspace():
            self.offset += 1
        if self.offset >= len(s):
            return None
        o, ch = self.offset, s[self.offset]
        if ch.isalpha() or ch == "`":
            t = self._identifier(o)
        elif ch.isdigit() or ch == ".":
            t = self._number(o)
        elif any(s.startswith(sym, o) for sym in SYMBOLS):
            sym = next(sym for sym in SYMBOLS if s.startswith(sym, o))
            t = Token(sym, "SYMBOL", o + len(sym))
        elif ch in "'\"":
            end = s.find(ch, o + 1)
            if end < 0:
                raise TokenizeException(f"unterminated string at offset {o}")
            t = Token(s[o + 1:end], "STRING", end + 1)
        else:
            raise TokenizeException(f"unexpected character {ch!r} at offset {o}")
        self.offset = t.endOffset
        return t

    def _identifier(self, o):
        s = self.sql
        if s[o] == "`":
            end = s.find("`", o + 1)
            if end < 0:
                raise TokenizeException(f"unterminated identifier at offset {o}")
            return Token(s[o + 1:end], "IDENTIFIER", end + 1)
        end = o
        while end < len(s) and (s[end].isalnum() or s[end] == "_"):
            end += 1
        text = s[o:end]
        return Token(text, "KEYWORD" if text.upper() in KEYWORDS else "IDENTIFIER", end)

    def _number(self, o):
        s, end = self.sql, o
        while end < len(s) and s[end].isdigit():
            end += 1
        is_float = end < len(s) and s[end] == "."
        if is_float:
            end += 1
            while end < len(s) and s[end].isdigit():
                end += 1
        return Token(s[o:end], "DOUBLE" if is_float else "LONG", end)


# ---- syntax tree ------------------------------------------------------------------------------------------------------------
class SqlExpr:
    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in self.__dict__.items())})"


class SqlIdentifier(SqlExpr):
    def __init__(self, id: str):
        self.id = id


class SqlFunction(SqlExpr):
    def __init__(self, id: str, args):
        self.id, self.args = id, list(args)


class SqlAlias(SqlExpr):
    def __init__(self, expr: SqlExpr, alias: SqlIdentifier):
        self.expr, self.alias = expr, alias


class SqlCast(SqlExpr):
    def __init__(self, expr: SqlExpr, dataType: SqlIdentifier):
        self.expr, self.dataType = expr, dataType


class SqlSort(SqlExpr):
    def __init__(self, expr: SqlExpr, asc: bool):
        self.expr, self.asc = expr, asc


class SqlLiteral(SqlExpr):                         # [+]
    def __init__(self, value):
        self.value = value


class SqlBinary(SqlExpr):                          # [+]
    def __init__(self, op: str, l: SqlExpr, r: SqlExpr):
        self.op, self.l, self.r = op, l, r


class SqlSelect(SqlExpr):
    def __init__(self, projection, selection, groupBy, orderBy, tableName: str):
        self.projection, self.selection, self.groupBy, self.orderBy, self.tableName = list(projection), selection, list(groupBy), list(orderBy), tableName


_BINARY = {"OR": (20, "OR"), "AND": (30, "AND"), "=": (40, "EQ"), "!=": (40, "NE"), "<>": (40, "NE"), "<": (40, "LT"), "<=": (40, "LE"),
           ">": (40, "GT"), ">=": (40, "GE"), "+": (50, "ADD"), "-": (50, "SUB"), "*": (60, "MUL"), "/": (60, "DIV")}


class SqlParser:
    """Pratt parser (Main.kt:1069-1205): AS binds at 10, a call's parenthesis at 70; the binary operators sit in between."""

    def __init__(self, tokens: TokenStream):
        self.tokens = tokens

    def parse(self, precedence: int = 0):
        expr = self.parsePrefix()
        if expr is None:
            return None
        while precedence < self.nextPrecedence():
            expr = self.parseInfix(expr)
        return expr

    def nextPrecedence(self) -> int:
        t = self.tokens.peek()
        if t is None:
            return 0
        if t.type == "KEYWORD":
            key = t.text.upper()
            return 10 if key == "AS" else _BINARY[key][0] if key in _BINARY else 0
        if t.type == "SYMBOL":
            return 70 if t.text == "(" else _BINARY[t.text][0] if t.text in _BINARY else 0
        return 0

    def parsePrefix(self):
        t = self.tokens.next()
        if t is None:
            return None
        if t.type == "KEYWORD":
            key = t.text.upper()
            if key == "SELECT":
                return self.parseSelect()
            if key == "CAST":
                return self.parseCast()
            if key in ("MAX", "MIN", "SUM", "COUNT", "DOUBLE"):
                return SqlIdentifier(t.text)
            if key == "DATE":                            # DATE 'yyyy-mm-dd' -> a Date32 literal (days since 1970-01-01)
                d = self.tokens.next()
                if d is None or d.type != "STRING":
                    raise IllegalStateException(f"Expected a date string after DATE, found {d}")
                import datetime
                try:
                    return SqlLiteral(datetime.date.fromisoformat(d.text))
                except ValueError:
                    raise SQLException(f"Invalid date '{d.text}'")
        elif t.type == "IDENTIFIER":
            return SqlIdentifier(t.text)
        elif t.type == "LONG":
            return SqlLiteral(int(t.text))
        elif t.type == "DOUBLE":
            return SqlLiteral(float(t.text))
        elif t.type == "STRING":
            return SqlLiteral(t.text)
        elif t.type == "SYMBOL" and t.text == "-":           # a negative number
            operand = self.parse(65)
            if isinstance(operand, SqlLiteral) and isinstance(operand.value, (int, float)):
                return SqlLiteral(-operand.value)
            raise IllegalStateException(f"Unexpected token {t}")
        elif t.type == "SYMBOL" and t.text == "(":
            inner = self.parse(0)
            if not self.tokens.consumeSymbol(")"):
                raise IllegalStateException(f"Expected ), found {self.tokens.peek()}")
            return inner
        raise IllegalStateException(f"Unexpected token {t}")

    def parseInfix(self, left):
        t = self.tokens.peek()
        if t.type == "KEYWORD" and t.text.upper() == "AS":
            self.tokens.next()
            return SqlAlias(left, self.parseIdentifier())
        if t.type == "SYMBOL" and t.text == "(":
            if not isinstance(left, SqlIdentifier):
                raise IllegalStateException("Unexpected LPAREN")
            self.tokens.next()
            args = self.parseExprList()
            if not self.tokens.consumeSymbol(")"):
                raise IllegalStateException(f"Expected ), found {self.tokens.peek()}")
            return SqlFunction(left.id, args)
        key = t.text.upper() if t.type == "KEYWORD" else t.text
        if key in _BINARY:
            self.tokens.next()
            prec, op = _BINARY[key]
            right = self.parse(prec)                   # left-associative
            if right is None:
                raise SQLException(f"Expected an expression after {t.text}, found EOF")
            return SqlBinary(op, left, right)
        raise IllegalStateException(f"Unexpected infix token {t}")

    def parseCast(self) -> SqlCast:
        if not self.tokens.consumeSymbol("("):
            raise IllegalStateException(f"Expected ( after CAST, found {self.tokens.peek()}")
        expr = self.parse(0)
        if not isinstance(expr, SqlAlias):
            raise SQLException(f"Expected CAST(expr AS type), found {expr}")
        if not self.tokens.consumeSymbol(")"):
            raise IllegalStateException(f"Expected ), found {self.tokens.peek()}")
        return SqlCast(expr.expr, expr.alias)

    def parseSelect(self) -> SqlSelect:
        projection = self.parseExprList()
        if not self.tokens.consumeKeyword("FROM"):
            raise IllegalStateException(f"Expected FROM keyword, found {self.tokens.peek()}")
        table = self.parse(0)
        if not isinstance(table, SqlIdentifier):
            raise SQLException(f"Expected a table name, found {table}")
        selection = None
        if self.tokens.consumeKeyword("WHERE"):
            selection = self.parse(0)
            if selection is None:
                raise SQLException("Expected a predicate after WHERE, found EOF")
        groupBy = self.parseExprList() if self.tokens.consumeKeywords(["GROUP", "BY"]) else []
        orderBy = self.parseOrder() if self.tokens.consumeKeywords(["ORDER", "BY"]) else []
        if self.tokens.peek() is not None:
            raise IllegalStateException(f"Unexpected token {self.tokens.peek()}")
        return SqlSelect(projection, selection, groupBy, orderBy, table.id)

    def parseOrder(self):
        out = []
        while True:
            e = self.parse(0)
            if e is None:
                break
            asc = True
            if self.tokens.consumeKeyword("DESC"):
                asc = False
            else:
                self.tokens.consumeKeyword("ASC")
            out.append(SqlSort(e, asc))
            if not self.tokens.consumeSymbol(","):
                break
        return out

    def parseExprList(self):
        out = []
        while True:
            e = self.parse(0)
            if e is None:
                break
            out.append(e)
            if not self.tokens.consumeSymbol(","):
                break
        return out

    def parseIdentifier(self) -> SqlIdentifier:
        e = self.parse(10)                              # the alias itself, not `alias AS ...`
        if e is None:
            raise SQLException("Expected identifier, found EOF")
        if not isinstance(e, SqlIdentifier):
            raise SQLException(f"Expected identifier, found {e}")
        return e


# ---- planner ------------------------------------------------------------------------------------------------------------------
_AGGREGATES = {"MAX": Max, "MIN": Min, "SUM": Sum, "COUNT": Count}


def parseDataType(id: str):
    """Main.kt:1284-1289 knows "double"; [+] bigint/long and string."""
    key = id.lower()
    if key == "double":
        return DoubleType
    if key in ("bigint", "long"):
        return Int64Type
    if key in ("string", "varchar"):
        return StringType
    raise SQLException(f"Invalid data type {id}")


def createLogicalExpr(expr: SqlExpr, input: DataFrame) -> LogicalExpr:
    """Main.kt:1264-1282."""
    if isinstance(expr, SqlIdentifier):
        return Column(expr.id)
    if isinstance(expr, SqlAlias):
        return Alias(createLogicalExpr(expr.expr, input), expr.alias.id)
    if isinstance(expr, SqlCast):
        return CastExpr(createLogicalExpr(expr.expr, input), parseDataType(expr.dataType.id))
    if isinstance(expr, SqlFunction):
        fn = _AGGREGATES.get(expr.id.upper())
        if fn is None or len(expr.args) != 1:
            raise SQLException(f"Invalid aggregate function: {expr.id}")
        return fn(createLogicalExpr(expr.args[0], input))
    if isinstance(expr, SqlLiteral):
        return lit(expr.value)
    if isinstance(expr, SqlBinary):
        return BinaryExpr(expr.op, createLogicalExpr(expr.l, input), createLogicalExpr(expr.r, input))
    raise SQLException(f"Cannot create logical expression from sql expression: {expr}")


def isAggregateExpr(expr: LogicalExpr) -> bool:
    return isinstance(expr, AggregateExpr) or (isinstance(expr, Alias) and isinstance(expr.expr, AggregateExpr))


def createDataFrame(select: SqlSelect, tables) -> DataFrame:
    """Main.kt:1218-1256: scan -> [selection] -> aggregate(group by, aggregates) -> projection that puts the select list in
    order over the aggregate's output by column index. [+] without aggregates: scan -> [selection] -> projection."""
    table = tables.get(select.tableName)
    if table is None:
        raise SQLException(f"No table named '{select.tableName}'")
    projectionExpr = [createLogicalExpr(e, table) for e in select.projection]
    naggr = sum(1 for e in projectionExpr if isAggregateExpr(e))
    if naggr == 0 and select.groupBy:
        raise SQLException("GROUP BY without aggregate expressions is not supported")
    plan = table
    if select.selection is not None:
        plan = plan.filter(createLogicalExpr(select.selection, table))
    if naggr == 0:
        return plan.project(projectionExpr)
    projection, aggrExpr = [], []
    numGroupCols, groupCount = len(select.groupBy), 0
    for e in projectionExpr:
        if isinstance(e, AggregateExpr):
            projection.append(ColumnIndex(numGroupCols + len(aggrExpr)))
            aggrExpr.append(e)
        elif isinstance(e, Alias) and isinstance(e.expr, AggregateExpr):
            projection.append(Alias(ColumnIndex(numGroupCols + len(aggrExpr)), e.alias))
            aggrExpr.append(e.expr)
        else:
            projection.append(ColumnIndex(groupCount))
            groupCount += 1
    groupByExpr = [createLogicalExpr(e, plan) for e in select.groupBy]
    return plan.aggregate(groupByExpr, aggrExpr).project(projection)
